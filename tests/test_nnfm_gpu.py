"""GPU parity of the fused nearest-neighbour feature matching against the reference composition (loss.py)."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


@pytest.fixture(params=['tcgen05', 'mma_sync'], autouse=True)
def nnfm_impl(request, cuda_lib):
    """Every test runs on both implementations behind nrf_nnfm_forward: tcgen05 + TMEM + bulk copies (nnfm_tc.cu, the
    default) and mma.sync (nnfm.cu)."""
    from nerfstyle_b200 import _lib
    _lib.lib().nrf_nnfm_set_mode(0 if request.param == 'tcgen05' else 1)
    yield request.param
    _lib.lib().nrf_nnfm_set_mode(0)


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


def test_full_size_implementations_agree(cuda_lib, dev):
    """BASELINE config 4 size (N1 = 11 844, N2 = 15 876, K = 768): the two implementations pick the same minima."""
    from nerfstyle_b200 import nnfm, _lib
    g = torch.Generator().manual_seed(4)
    N1, N2, K = 11844, 15876, 768
    a = torch.randn(N1, K, generator=g).to(dev)
    b = torch.randn(N2, K, generator=g).to(dev)
    a_hat = a / a.norm(dim=1, keepdim=True)
    b_hat = b / b.norm(dim=1, keepdim=True)
    preds = torch.randint(0, 8, (N1,), generator=g).to(dev)
    clusters = torch.randint(0, 8, (N2,), generator=g).to(dev)
    match = list(range(8))
    res = []
    for mode in (0, 1):
        _lib.lib().nrf_nnfm_set_mode(mode)
        res.append(nnfm.nn_match(a_hat, b_hat, preds, clusters, match))
    _lib.lib().nrf_nnfm_set_mode(0)
    torch.testing.assert_close(res[0][0], res[1][0], rtol=0, atol=1e-4)
    assert float((res[0][1] == res[1][1]).float().mean()) > 0.999
    # and against a dense fp32 evaluation of a row sample
    idx = torch.arange(0, N1, 97, device=dev)
    d = 1.0 - a_hat[idx].half().float() @ b_hat.half().float().T
    d[preds[idx][:, None] != clusters[None, :]] = float('inf')
    torch.testing.assert_close(res[0][0][idx], d.min(dim=1).values, rtol=0, atol=2e-4)
