"""The segment-wise matching loss against outputs of the REFERENCE's own loss.py (tests/golden/matching.npz, written by
tests/golden/make_matching_golden.py from the unmodified `SemanticStyleLoss.init_feats / update_matching / forward`,
`labels_downscale`, `cosine_dists`, `NNFMStyleLoss.forward`): the CPU oracle restatement (oracle/matching.py) on the
CPU, and the tensor-core kernel behind nerfstyle_b200.nnfm.SemanticStyleLoss on the GPU."""
import os

import numpy as np
import pytest
import torch

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), 'golden', 'matching.npz')


def _cases():
    g = np.load(GOLD)
    for name in g['names']:
        name = str(name)
        yield name, {k.split('/', 1)[1]: g[k] for k in g.files if k.startswith(name + '/')}


def test_oracle_matching_equals_reference_outputs():
    from oracle import matching as om
    n = 0
    for name, c in _cases():
        img, sty = torch.from_numpy(c['image_feat']), torch.from_numpy(c['style_feat'])
        preds, clusters = torch.from_numpy(c['preds']), torch.from_numpy(c['clusters_full'])
        matching = None if c['matching_in'][0] < 0 else c['matching_in'].tolist()
        n_cls = len(c['matching'])
        loss, md, arg, ps, cs, m = om.semantic_style_loss(img, sty, preds, clusters, matching, n_cls)
        assert np.array_equal(ps.numpy(), c['preds_small']), name               # labels_downscale: integers, exact
        assert np.array_equal(cs.numpy(), c['clusters_small']), name
        assert list(m) == c['matching'].tolist(), name                          # Hungarian assignment
        assert np.array_equal(np.isfinite(md.numpy()), np.isfinite(c['min_dists'])), name
        fin = np.isfinite(c['min_dists'])
        np.testing.assert_allclose(md.numpy()[fin], c['min_dists'][fin], rtol=0, atol=1e-6)
        assert np.array_equal(arg.numpy()[fin], c['argmin'][fin]), name
        if np.isfinite(c['loss']):
            assert abs(float(loss) - float(c['loss'])) < 1e-6, name
        else:
            assert np.isinf(float(loss)), name
        assert abs(float(om.nnfm_loss(img, sty)) - float(c['nnfm_loss'])) < 1e-6, name
        n += 1
    assert n == 5


@pytest.mark.gpu
@pytest.mark.parametrize('mode', [0, 1])
def test_product_semantic_style_loss_equals_reference_outputs(cuda_lib, dev, mode):
    """nerfstyle_b200.nnfm.SemanticStyleLoss / NNFMStyleLoss (tcgen05 kernel, mode 0; mma.sync, mode 1) reproduce the
    reference's loss value on the reference's inputs: fp16 tensor-core operands, so 2e-4 absolute on a cosine distance."""
    from nerfstyle_b200 import nnfm, _lib
    _lib.lib().nrf_nnfm_set_mode(mode)
    try:
        for name, c in _cases():
            img = torch.from_numpy(c['image_feat']).to(dev).requires_grad_(True)
            sty = torch.from_numpy(c['style_feat']).to(dev)
            preds = torch.from_numpy(c['preds']).to(dev)
            matching = None if c['matching_in'][0] < 0 else c['matching_in'].tolist()
            n_cls = len(c['matching'])
            obj = nnfm.SemanticStyleLoss(['relu3_1'], None, matching, clusters=torch.from_numpy(c['clusters_full']))
            if name == 'empty_cluster':
                obj.n_clusters = n_cls
            obj.init_feats({'relu3_1': sty[None]}, n_cls)
            assert np.array_equal(obj.clusters.cpu().numpy(), c['clusters_small'])
            loss = obj.forward({'relu3_1': img[None]}, None, preds, 0)
            assert [int(v) for v in obj.matching] == c['matching'].tolist(), name
            assert np.array_equal(nnfm.labels_downscale(preds, img.shape[-2:]).cpu().numpy(), c['preds_small'])
            if np.isfinite(c['loss']):
                assert abs(float(loss) - float(c['loss'])) < 2e-4, (name, float(loss), float(c['loss']))
                loss.backward()
                assert torch.isfinite(img.grad).all() and float(img.grad.abs().max()) > 0
            else:
                assert np.isinf(float(loss)), name
            # per-row minima / arg-minima through the kernel entry point
            C = img.shape[0]
            a = img.detach().reshape(C, -1).t()
            b = sty.reshape(C, -1).t()
            md, am = nnfm.nn_match(a / a.norm(dim=1, keepdim=True), b / b.norm(dim=1, keepdim=True),
                                   torch.from_numpy(c['preds_small']).reshape(-1).to(dev), obj.clusters.reshape(-1),
                                   c['matching'].tolist())
            fin = np.isfinite(c['min_dists'])
            assert np.array_equal(np.isfinite(md.cpu().numpy()), fin), name
            np.testing.assert_allclose(md.cpu().numpy()[fin], c['min_dists'][fin], rtol=0, atol=2e-4)
            assert (am.cpu().numpy()[fin] == c['argmin'][fin]).mean() > 0.97, name          # fp16 near-ties may flip
            assert bool((am.cpu().numpy()[~fin] == -1).all())
            nl = nnfm.NNFMStyleLoss(['relu3_1'])({'relu3_1': img.detach()[None]}, {'relu3_1': sty[None]})
            assert abs(float(nl) - float(c['nnfm_loss'])) < 2e-4, name
    finally:
        _lib.lib().nrf_nnfm_set_mode(0)
