"""GPU parity of the fused nearest-neighbour feature matching against the reference composition (loss.py)."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize('N1,N2,K,masked', [(1000, 1300, 768, True), (257, 129, 64, False), (1, 5, 8, True), (640, 2048, 256, True)])
def test_nn_match(cuda_lib, dev, N1, N2, K, masked):
    from nerfstyle_b200 import nnfm
    from oracle import matching as om
    g = torch.Generator().manual_seed(N1)
    a = torch.randn(N1, K, generator=g).to(dev)
    b = torch.randn(N2, K, generator=g).to(dev)
    n_class = 8
    preds = torch.randint(-1, n_class + 1, (N1,), generator=g).to(dev)       # includes classes outside [0, n_class)
    clusters = torch.randint(0, n_class, (N2,), generator=g).to(dev)
    if N2 > 100:
        clusters[clusters == 5] = 4                                          # cluster 5 is empty: rows matched to it get +inf
    match = torch.randperm(n_class, generator=g).tolist()
    a_hat = a / a.norm(dim=1, keepdim=True)
    b_hat = b / b.norm(dim=1, keepdim=True)
    md, am = nnfm.nn_match(a_hat, b_hat, preds if masked else None, clusters if masked else None, match if masked else None)
    # reference composition on the same fp16-rounded operands (the tensor cores see fp16)
    loss, emd, eam = om.semantic_nn_loss(a_hat.half().float(), b_hat.half().float(), preds if masked else None,
                                         clusters if masked else None, match if masked else None, n_class)
    # cosine_dists re-normalises; compare against plain 1 - dot of the rounded operands instead
    d = 1.0 - a_hat.half().float() @ b_hat.half().float().T
    if masked:
        for i in range(n_class):
            d[(preds == i)[:, None] & (clusters != match[i])[None, :]] = float('inf')
    emd, eam = torch.min(d, dim=1)
    fin = torch.isfinite(emd)
    assert torch.equal(torch.isfinite(md), fin)
    torch.testing.assert_close(md[fin], emd[fin], rtol=0, atol=2e-4)
    same = (am.long() == eam)[fin]
    # different accumulation order can flip near-ties: the chosen column must be (nearly) as good
    chosen = d[torch.arange(N1, device=dev)[fin], am.long()[fin]]
    assert float((chosen - emd[fin]).abs().max()) <= 2e-4 and float(same.float().mean()) > 0.99
    assert bool((am[~fin] == -1).all())


def test_semantic_loss_value_and_grad(cuda_lib, dev):
    from nerfstyle_b200 import nnfm
    from oracle import matching as om
    g = torch.Generator().manual_seed(0)
    C, h, w, hs, ws_ = 96, 20, 24, 28, 30
    img = torch.randn(C, h, w, generator=g).to(dev).requires_grad_(True)
    sty = torch.randn(C, hs, ws_, generator=g).to(dev)
    preds = torch.randint(0, 4, (h, w), generator=g).to(dev)
    clusters = torch.randint(0, 4, (hs, ws_), generator=g).to(dev)
    match = [2, 0, 3, 1]
    loss = nnfm.semantic_nnfm_loss(img, sty, preds, clusters, match)
    loss.backward()
    img2 = img.detach().clone().requires_grad_(True)
    eloss, _, _ = om.semantic_nn_loss(img2.reshape(C, -1).t(), sty.reshape(C, -1).t(), preds, clusters, match, 4)
    eloss.backward()
    assert abs(float(loss) - float(eloss)) < 2e-4
    # gradients agree wherever the fp16 arg-min equals the fp32 arg-min (all but near-ties)
    diff = (img.grad - img2.grad).reshape(C, -1).abs().amax(dim=0)
    assert float((diff < 1e-6).float().mean()) > 0.97
