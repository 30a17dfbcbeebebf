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
