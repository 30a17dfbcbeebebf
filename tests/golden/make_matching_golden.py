"""Generates tests/golden/matching.npz by running the REFERENCE's own `loss.py` -- `SemanticStyleLoss.init_feats`,
`.update_matching`, `.forward` (loss.py:115-214), `labels_downscale` (:23-28), `cosine_dists` (:32-36) and
`NNFMStyleLoss.forward` (:93-112) -- unmodified, on the CPU of this container.

loss.py imports matplotlib.pyplot at module scope (only used by a debug helper); an empty stand-in module is registered
for it.  `SemanticStyleLoss.__init__` moves the cluster map to the GPU with `.cuda()` when given a file path, so the
object is constructed with `clusters_path=None` and the four attributes that branch would have set (`use_matching`,
`clusters`, `n_clusters`, `matching`) are assigned directly; everything after construction is the reference's code.
These are outputs of the reference itself: the matching loss is PINNED by the reference.
Run from the repo root:  python tests/golden/make_matching_golden.py
"""
import os
import sys
import types

import numpy as np
import torch

REF = '/root/reference'
HERE = os.path.dirname(os.path.abspath(__file__))


def load_reference_loss():
    if 'matplotlib' not in sys.modules:
        mpl = types.ModuleType('matplotlib')
        plt = types.ModuleType('matplotlib.pyplot')
        mpl.pyplot = plt
        sys.modules['matplotlib'] = mpl
        sys.modules['matplotlib.pyplot'] = plt
    sys.path.insert(0, REF)
    import loss as ref_loss
    assert ref_loss.__file__.startswith(REF)
    return ref_loss


def smooth_labels(h, w, n, seed):
    """Blocky label map [h, w] with values in [0, n) (a few pixels set to n: a class the loss has no entry for)."""
    rs = np.random.RandomState(seed)
    coarse = rs.randint(0, n, size=(max(h // 6, 1) + 1, max(w // 6, 1) + 1))
    lab = coarse[np.arange(h)[:, None] // 6, np.arange(w)[None, :] // 6]
    return lab.astype(np.int64)


CASES = [
    # name, C, image-feature size, style-feature size, full-res prediction size, full-res cluster size, classes, matching
    ('given_matching', 48, (12, 16), (14, 14), (47, 63), (56, 56), 4, [2, 0, 3, 1]),
    ('hungarian', 32, (9, 13), (11, 12), (36, 50), (44, 48), 3, None),
    ('identity_8', 64, (10, 14), (12, 12), (40, 56), (48, 48), 8, list(range(8))),
    # some pixels carry label n_cls (no mask entry: every style column stays valid for them, loss.py:205-209)
    ('unlisted_class', 32, (12, 16), (14, 14), (47, 63), (56, 56), 4, [1, 0, 3, 2]),
    # cluster 3 does not occur in the style segmentation: rows matched to it have no valid column, the loss is +inf
    ('empty_cluster', 32, (12, 16), (14, 14), (47, 63), (56, 56), 4, [3, 0, 2, 1]),
]


def main():
    L = load_reference_loss()
    out = {}
    names = []
    for name, C, (h, w), (hs, ws), (H, W), (Hs, Ws), n_cls, matching in CASES:
        g = torch.Generator().manual_seed(len(name) * 7 + C)
        image_feat = torch.randn(C, h, w, generator=g)
        style_feat = torch.randn(C, hs, ws, generator=g)
        preds = torch.from_numpy(smooth_labels(H, W, n_cls, 1))
        clusters_full = smooth_labels(Hs, Ws, n_cls, 2)
        if name == 'unlisted_class':
            preds[::5, ::3] = n_cls
        if name == 'empty_cluster':
            clusters_full[clusters_full == 3] = 2
        obj = L.SemanticStyleLoss(['relu3_1'], None, None)
        obj.use_matching = True
        obj.clusters = torch.tensor(clusters_full)
        obj.n_clusters = n_cls
        obj.matching = None if matching is None else list(matching)
        obj.init_feats({'relu3_1': style_feat[None]}, n_cls)            # loss.py:147-170
        import io
        import contextlib
        with contextlib.redirect_stdout(io.StringIO()):                 # update_matching prints the assignment
            loss = obj.forward({'relu3_1': image_feat[None]}, None, preds, 0)    # loss.py:187-214
        # by-products, through the reference's own helpers
        preds_small = L.labels_downscale(preds, (h, w))
        a = image_feat.reshape(C, -1).t()
        b = style_feat.reshape(C, -1).t()
        d = L.cosine_dists(a, b)
        for i in range(n_cls):
            im = (preds_small == i).reshape(-1)
            sm = (obj.clusters != obj.matching[i]).reshape(-1)
            d[im[:, None] & sm[None, :]] = float('inf')
        min_d, arg = torch.min(d, dim=1)
        assert float(torch.mean(min_d)) == float(loss)
        nn = L.NNFMStyleLoss(['relu3_1'])
        nn_loss = nn.forward({'relu3_1': image_feat[None]}, {'relu3_1': style_feat[None]})   # loss.py:93-112
        names.append(name)
        out[name + '/image_feat'] = image_feat.numpy()
        out[name + '/style_feat'] = style_feat.numpy()
        out[name + '/preds'] = preds.numpy()
        out[name + '/clusters_full'] = clusters_full
        out[name + '/clusters_small'] = obj.clusters.numpy()
        out[name + '/preds_small'] = preds_small.numpy()
        out[name + '/matching_in'] = np.array([-1] if matching is None else matching, np.int64)
        out[name + '/matching'] = np.asarray(obj.matching, np.int64)
        out[name + '/style_feats_mean'] = obj.style_feats_mean.numpy()
        out[name + '/style_centroids'] = obj.style_centroids.numpy()
        out[name + '/loss'] = np.float64(float(loss))
        out[name + '/min_dists'] = min_d.numpy()
        out[name + '/argmin'] = arg.numpy()
        out[name + '/nnfm_loss'] = np.float64(float(nn_loss))
        print(name, 'loss', float(loss), 'finite rows', int(torch.isfinite(min_d).sum()), '/', min_d.numel(), 'matching', list(obj.matching),
              'nnfm', float(nn_loss))
    out['names'] = np.array(names)
    np.savez_compressed(os.path.join(HERE, 'matching.npz'), **out)
    print('wrote', os.path.join(HERE, 'matching.npz'))


if __name__ == '__main__':
    main()
