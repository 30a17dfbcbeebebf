"""Nearest-neighbour feature-matching loss of the stylization stage (loss.py:32-36, 187-214) on the fused
tensor-core kernel (csrc/nnfm.cu): cosine GEMM + class/cluster mask + row arg-min without materialising the
N1 x N2 distance matrix.

    loss = semantic_nnfm_loss(image_feat [C,h,w], style_feat [C,hs,ws], preds_small [h,w], clusters [hs,ws], matching)

equals the reference's `SemanticStyleLoss.forward` value (given `matching`); the gradient w.r.t. image_feat is
the exact gradient of the selected minimum (d/da of 1 - a_hat . b_hat[j*]), obtained by re-evaluating the N1
selected dot products differentiably in fp32 -- the arg-min itself is piecewise constant.
"""
import torch

from . import _lib as L


@torch.no_grad()
def nn_match(a_hat, b_hat, a_label=None, b_label=None, match=None):
    """a_hat [N1,K], b_hat [N2,K]: L2-normalised rows (any float dtype; fp16 copies are made for the tensor cores).
    Returns (min_dist [N1] f32, argmin [N1] i32); rows with no allowed column get (+inf, -1)."""
    L.require_cuda(a_hat, b_hat)
    a16 = a_hat.to(torch.float16).contiguous()
    b16 = b_hat.to(torch.float16).contiguous()
    N1, K = a16.shape
    N2 = b16.shape[0]
    dev = a16.device
    min_dist = torch.empty(N1, dtype=torch.float32, device=dev)
    argmin = torch.empty(N1, dtype=torch.int32, device=dev)
    lib = L.lib()
    scratch = torch.empty(int(lib.nrf_nnfm_scratch_bytes(N1, N2, K)), dtype=torch.uint8, device=dev)
    n_class = 0
    if match is not None:
        a_label = a_label.to(device=dev, dtype=torch.int32).contiguous()
        b_label = b_label.to(device=dev, dtype=torch.int32).contiguous()
        match = torch.as_tensor(match, dtype=torch.int32, device=dev).contiguous()
        n_class = match.numel()
    with torch.cuda.device(dev):
        L.check(lib.nrf_nnfm_forward(L.ptr(a16), L.ptr(b16), N1, N2, K, L.ptr(a_label) if match is not None else None,
                                     L.ptr(b_label) if match is not None else None, L.ptr(match), n_class,
                                     L.ptr(min_dist), L.ptr(argmin), L.ptr(scratch), L.stream_of(a16)), 'nnfm_forward')
    return min_dist, argmin


def _normalize_rows(f):
    return f / torch.linalg.norm(f, dim=1)[:, None]       # loss.py:33-34


def semantic_nnfm_loss(image_feat, style_feat, preds_small=None, clusters=None, matching=None):
    """loss.py:187-214 with the matrix-free kernel.  image_feat [C,h,w], style_feat [C,hs,ws]."""
    C = image_feat.shape[0]
    a = image_feat.reshape(C, -1).t().float()             # 'c h w -> (h w) c'
    b = style_feat.reshape(C, -1).t().float()
    a_hat = _normalize_rows(a)
    b_hat = _normalize_rows(b.detach())
    if matching is not None:
        _, j = nn_match(a_hat.detach(), b_hat, preds_small.reshape(-1), clusters.reshape(-1), matching)
    else:
        _, j = nn_match(a_hat.detach(), b_hat)
    valid = j >= 0
    jj = j.clamp(min=0).long()
    sim = (a_hat * b_hat[jj]).sum(dim=1)
    dist = torch.where(valid, 1.0 - sim, torch.full_like(sim, float('inf')))
    return dist.mean()


def labels_downscale(labels, new_dim):
    """loss.py:23-28: the class map at the feature-map resolution (rows / columns linspace(0, H-1, NH) truncated)."""
    H, W = labels.shape
    NH, NW = new_dim
    r_indices = torch.linspace(0, H - 1, NH).long().to(labels.device)
    c_indices = torch.linspace(0, W - 1, NW).long().to(labels.device)
    return labels[r_indices[:, None], c_indices]


class NNFMStyleLoss(torch.nn.Module):
    """loss.py:93-112 on the fused kernel: mean over image positions of the nearest style feature's cosine distance."""

    def __init__(self, keys):
        super().__init__()
        self.keys = keys

    def forward(self, feats1, feats2):
        loss = 0
        for k in self.keys:
            loss = loss + semantic_nnfm_loss(feats1[k].squeeze(0), feats2[k].squeeze(0))
        return loss


class SemanticStyleLoss(torch.nn.Module):
    """loss.py:115-214 with the same constructor / init_feats / forward interface; the N1 x N2 distance matrix, its
    per-class `inf` masking loop and the row minimum are one tensor-core kernel (nn_match).  `clusters` may be given as
    a tensor instead of a `.npz` path."""

    def __init__(self, keys, clusters_path=None, matching=None, clusters=None):
        super().__init__()
        self.keys = keys
        self.ready = False
        self.clusters = None
        self.matching = None
        self.use_matching = False
        if clusters_path is not None or clusters is not None:
            import numpy as np
            self.use_matching = True
            seg = np.load(str(clusters_path))['seg_map'] if clusters is None else np.asarray(torch.as_tensor(clusters).cpu())
            ids = np.unique(seg)
            if ids[0] < 0:
                ids = ids[1:]
            self.n_clusters = len(ids) if matching is None else max(len(ids), int(max(matching)) + 1)
            self.clusters = torch.as_tensor(seg)
            self.matching = matching

    @torch.no_grad()
    def init_feats(self, all_style_feats, num_classes):
        style_feats = all_style_feats[self.keys[0]].squeeze(0)
        self.style_feats = style_feats
        if self.use_matching:
            import torch.nn.functional as F
            self.clusters = F.interpolate(self.clusters.to(style_feats.device)[None, None].float(), style_feats.shape[1:])
            self.clusters = self.clusters[0, 0].to(torch.long)
            self.num_classes = num_classes
        self.ready = True

    def update_matching(self, image_feats, preds):
        """loss.py:172-185 (host-side Hungarian assignment on K x K means; not a hot path)."""
        import numpy as np
        from scipy.optimize import linear_sum_assignment

        def centroid(mask):
            H, W = mask.shape
            N = torch.sum(mask)
            r = torch.sum(torch.sum(mask, dim=1) * torch.arange(H, device=mask.device)) / N / H
            c = torch.sum(torch.sum(mask, dim=0) * torch.arange(W, device=mask.device)) / N / W
            return torch.stack((r, c))
        preds_small = labels_downscale(preds, image_feats.shape[-2:])
        im_mean = torch.stack([torch.mean(image_feats[:, preds_small == i], dim=1) for i in range(self.num_classes)])
        im_cent = torch.stack([centroid(preds == i) for i in range(self.num_classes)])
        st_mean = torch.stack([torch.mean(self.style_feats[:, self.clusters == i], dim=1) for i in range(self.n_clusters)])
        st_cent = torch.stack([centroid(self.clusters == i) for i in range(self.n_clusters)])
        feat_d = 1.0 - _normalize_rows(im_mean.float()) @ _normalize_rows(st_mean.float()).T
        cost = feat_d + torch.linalg.norm(im_cent[:, None] - st_cent[None], dim=-1)
        self.matching = linear_sum_assignment(np.nan_to_num(cost.detach().cpu().numpy()))[1]

    def forward(self, feats1, _, preds, iter=0):
        assert self.ready
        image_feat = feats1[self.keys[0]].squeeze(0)
        if not self.use_matching:
            return semantic_nnfm_loss(image_feat, self.style_feats)
        if self.matching is None:
            self.update_matching(image_feat.detach(), preds)
        preds_small = labels_downscale(preds, image_feat.shape[-2:])
        # classes outside [0, num_classes) have no mask entry in the reference: label them -1 (= unmasked in the kernel)
        ps = torch.where((preds_small >= 0) & (preds_small < self.num_classes), preds_small, torch.full_like(preds_small, -1))
        return semantic_nnfm_loss(image_feat, self.style_feats, ps, self.clusters, [int(m) for m in self.matching])
