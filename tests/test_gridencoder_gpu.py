"""GPU parity of the hash-grid encoder against the CPU oracle.  fp32 forward: bit-exact (same hash rows, same fma
chain).  fp16 tables: <= 1 half ulp.  Backward (sum order differs from any atomics schedule): rel 1e-5 of max."""
import os

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(__file__), 'golden')


def _default_encoder(dev, seed=0, std=1.0):
    from nerfstyle_b200.model import get_grid_encoder
    enc = get_grid_encoder(max_bound=4.0).to(dev)
    g = torch.Generator().manual_seed(seed)
    with torch.no_grad():
        enc.embeddings.copy_(((torch.rand(enc.embeddings.shape, generator=g) * 2 - 1) * std).to(dev))
    return enc


def _points(B, seed, dev, lo=0.0, hi=1.0):
    g = torch.Generator().manual_seed(seed)
    return (torch.rand(B, 3, generator=g) * (hi - lo) + lo).to(dev)


@pytest.mark.parametrize('B', [1, 255, 4099])
@pytest.mark.parametrize('lpt', [16, 4, 1])
def test_forward_fp32_bit_exact(cuda_lib, oracle, dev, B, lpt):
    cuda_lib.nrf_grid_set_tuning(lpt, 0, -1)
    try:
        enc = _default_encoder(dev)
        x = _points(B, B, dev, -1.0, 1.0)                       # GridEncoder.forward remaps [-1,1] -> [0,1]
        x[0] = torch.tensor([1.0, -1.0, 0.0], device=dev)       # exact borders
        out = enc(x)
        inp = ((x + 1) / 2).cpu().numpy()
        eo, _ = oracle.grid_encode_forward(inp, enc.embeddings.detach().cpu().numpy(), enc.offsets.cpu().numpy(),
                                           enc.per_level_scale, 16, False, 0, True, 0)
        assert out.shape == (B, 32) and out.dtype == torch.float32
        assert np.array_equal(out.detach().cpu().numpy().view(np.uint32), eo.view(np.uint32))
    finally:
        cuda_lib.nrf_grid_set_tuning(16, 0, -1)


def test_forward_golden_and_oob(cuda_lib, dev):
    from nerfstyle_b200.gridencoder import grid_encode
    g = np.load(os.path.join(GOLD, 'grid_small.npz'))
    offs, pls = torch.from_numpy(g['offsets']).to(dev), float(g['per_level_scale'])
    emb = torch.from_numpy(np.random.RandomState(int(g['emb_seed'])).uniform(-1, 1, (int(g['offsets'][-1]), 2)).astype(np.float32)).to(dev)
    x = torch.from_numpy(g['inputs']).to(dev)
    out = grid_encode(x, emb, offs, pls, 16, False, 0, True, 0)
    assert np.array_equal(out.cpu().numpy().view(np.uint32), g['outputs'].view(np.uint32))
    assert float(out[3].abs().sum()) == 0.0                     # out-of-range input -> zeros (gridencoder.cu:107-132)
    with torch.autocast('cuda', dtype=torch.float16):
        out_h = grid_encode(x, emb, offs, pls, 16, False, 0, True, 0)
    assert out_h.dtype == torch.float16
    diff = np.abs(out_h.float().cpu().numpy() - g['outputs_half'].astype(np.float32))
    ulp = np.maximum(np.abs(g['outputs_half'].astype(np.float32)), 2.0 ** -14) * 2.0 ** -10
    assert (diff <= ulp + 1e-12).all()


def test_forward_fp16_autocast(cuda_lib, oracle, dev):
    enc = _default_encoder(dev)
    x = _points(3001, 7, dev, -1.0, 1.0)
    with torch.autocast('cuda', dtype=torch.float16):
        out = enc(x)
    assert out.dtype == torch.float16
    inp = ((x + 1) / 2).cpu().numpy()
    eo, _ = oracle.grid_encode_forward(inp, enc.embeddings.detach().half().cpu().numpy(), enc.offsets.cpu().numpy(),
                                       enc.per_level_scale, 16, False, 0, True, 0, half=True)
    a, b = out.detach().float().cpu().numpy(), eo.astype(np.float32)
    ulp = np.maximum(np.abs(b), 2.0 ** -14) * 2.0 ** -10
    assert (np.abs(a - b) <= ulp + 1e-12).all()
    assert (a != b).mean() < 0.02


@pytest.mark.parametrize('agg', [0, 1])
@pytest.mark.parametrize('lpt', [16, 4])
def test_backward_fp32(cuda_lib, oracle, dev, agg, lpt):
    cuda_lib.nrf_grid_set_tuning(0, lpt, agg)
    try:
        enc = _default_encoder(dev)
        B = 6000
        # half of the points along a few rays (coherent -> exercises the warp aggregation), half random
        t = torch.linspace(0, 1, B // 2, device=dev)[:, None]
        line = torch.tensor([[-0.9, -0.7, 0.3]], device=dev) + t * torch.tensor([[1.7, 1.1, 0.4]], device=dev)
        x = torch.cat([line, _points(B - B // 2, 3, dev, -1.0, 1.0)], dim=0)
        x[5] = 3.0                                             # one out-of-range point: no gradient
        out = enc(x)
        g = torch.Generator().manual_seed(1)
        grad = torch.randn(out.shape, generator=g).to(dev)
        out.backward(grad)
        ge = enc.embeddings.grad.cpu().numpy()
        inp = ((x + 1) / 2).cpu().numpy()
        ege = oracle.grid_encode_backward(grad.cpu().numpy(), inp, enc.offsets.cpu().numpy(), enc.embeddings.shape[0], 2,
                                          enc.per_level_scale, 16, 0, True, 0)
        scale = np.abs(ege).max()
        assert np.abs(ge - ege).max() <= 1e-5 * scale
        assert (ge != 0).sum() == (ege != 0).sum()
    finally:
        cuda_lib.nrf_grid_set_tuning(0, 16, 1)


def test_backward_fp16_autocast(cuda_lib, oracle, dev):
    """fp16 grads are accumulated into an fp32 table gradient (more accurate than the reference's half atomics)."""
    enc = _default_encoder(dev)
    x = _points(4000, 11, dev, -1.0, 1.0)
    with torch.autocast('cuda', dtype=torch.float16):
        out = enc(x)
    g = torch.Generator().manual_seed(2)
    grad = torch.randn(out.shape, generator=g).to(dev).half()
    out.backward(grad)
    ge = enc.embeddings.grad
    assert ge.dtype == torch.float32
    ege = oracle.grid_encode_backward(grad.float().cpu().numpy(), ((x + 1) / 2).cpu().numpy(), enc.offsets.cpu().numpy(),
                                      enc.embeddings.shape[0], 2, enc.per_level_scale, 16, 0, True, 0)
    assert np.abs(ge.cpu().numpy() - ege).max() <= 1e-5 * np.abs(ege).max()


def test_backward_fp16_reference_mode(cuda_lib, oracle, dev):
    """grad_table_dtype = f16 reproduces the reference's __half2 atomics (every add rounds to half)."""
    from nerfstyle_b200 import _lib as L
    enc = _default_encoder(dev)
    B = 4000
    x = _points(B, 11, dev, 0.0, 1.0).contiguous()
    g = torch.Generator().manual_seed(2)
    grad = torch.randn(B, 32, generator=g).to(dev).half().contiguous()
    ge = torch.zeros(enc.embeddings.shape, dtype=torch.float16, device=dev)
    S = float(np.float32(np.log2(enc.per_level_scale)))
    L.check(cuda_lib.nrf_grid_encode_backward(grad.data_ptr(), x.data_ptr(), None, enc.offsets.data_ptr(), ge.data_ptr(), B, 3, 2,
                                              16, S, 16, 0, None, None, 0, 1, 0, L.DTYPE_F16, L.DTYPE_F16, 1,
                                              torch.cuda.current_stream().cuda_stream), 'bwd')
    ege = oracle.grid_encode_backward(grad.float().cpu().numpy(), x.cpu().numpy(), enc.offsets.cpu().numpy(),
                                      enc.embeddings.shape[0], 2, enc.per_level_scale, 16, 0, True, 0)
    err = np.abs(ge.float().cpu().numpy() - ege)
    assert err.max() <= 8 * 2.0 ** -10 * np.abs(ege).max()
    assert np.median(err[ege != 0] / np.abs(ege[ege != 0])) < 2e-3


@pytest.mark.parametrize('D,C,gridtype,align', [(3, 4, 0, False), (2, 2, 1, True), (3, 1, 0, True), (3, 8, 1, False)])
def test_generic_paths_and_input_grads(cuda_lib, oracle, dev, D, C, gridtype, align):
    from nerfstyle_b200.gridencoder import GridEncoder
    enc = GridEncoder(input_dim=D, num_levels=6, level_dim=C, per_level_scale=1.7, base_resolution=8, log2_hashmap_size=12,
                      gridtype='hash' if gridtype == 0 else 'tiled', align_corners=align).to(dev)
    g = torch.Generator().manual_seed(4)
    with torch.no_grad():
        enc.embeddings.copy_((torch.rand(enc.embeddings.shape, generator=g) * 2 - 1).to(dev))
    x = (torch.rand(1500, D, generator=g) * 2 - 1).to(dev).requires_grad_(True)
    out = enc(x)
    grad = torch.randn(out.shape, generator=g).to(dev)
    out.backward(grad)
    inp = ((x.detach() + 1) / 2).cpu().numpy()
    eo, edy = oracle.grid_encode_forward(inp, enc.embeddings.detach().cpu().numpy(), enc.offsets.cpu().numpy(), 1.7, 8, True,
                                         gridtype, align, 0)
    assert np.array_equal(out.detach().cpu().numpy().view(np.uint32), eo.view(np.uint32))
    ege = oracle.grid_encode_backward(grad.cpu().numpy(), inp, enc.offsets.cpu().numpy(), enc.embeddings.shape[0], C, 1.7, 8,
                                      gridtype, align, 0)
    assert np.abs(enc.embeddings.grad.cpu().numpy() - ege).max() <= 1e-5 * np.abs(ege).max()
    egi = oracle.grid_input_backward(grad.cpu().numpy(), edy, 1500, D, C, 6) * 0.5     # d/dx of (x+1)/2
    np.testing.assert_allclose(x.grad.cpu().numpy(), egi, rtol=1e-5, atol=1e-5 * np.abs(egi).max())


def test_empty_and_errors(cuda_lib, dev):
    enc = _default_encoder(dev)
    out = enc(torch.zeros(0, 3, device=dev))
    assert out.shape == (0, 32)
    with pytest.raises(RuntimeError):
        enc(torch.zeros(4, 3))                                   # CPU tensor: no fallback
    assert cuda_lib.nrf_grid_encode_forward(1, 1, 1, 1, 4, 7, 2, 16, 0.5, 16, 0, None, 0, 1, 0, 0, 1, None) == -2


@pytest.mark.parametrize('amp', [False, True])
def test_dual_equals_two_single_passes(cuda_lib, dev, amp):
    """nrf_grid_encode_*_dual (one index computation for two tables) == two independent encoder calls: forward bit for
    bit, backward up to the float-atomics summation order.  Points mimic
    ray samples (runs in the same cell) plus out-of-range rows."""
    from nerfstyle_b200.gridencoder import grid_encode_dual, same_geometry
    ea, eb = _default_encoder(dev, 1), _default_encoder(dev, 2)
    assert same_geometry(ea, eb)
    B = 20011
    g = torch.Generator().manual_seed(5)
    base = torch.rand(B // 32 + 1, 3, generator=g).repeat_interleave(32, dim=0)[:B]
    x = (base + torch.arange(B)[:, None] % 32 * 4e-4).to(dev) * 2 - 1           # short rays of 32 samples
    x[7] = 3.0                                                                   # out of range -> zero row, no gradient
    if True:
        res = []
        for dual in (True, False):
            for e in (ea, eb):
                e.embeddings.grad = None
            with torch.autocast('cuda', dtype=torch.float16, enabled=amp):
                if dual:
                    oa, ob = grid_encode_dual(x, ea, eb)
                else:
                    oa, ob = ea(x), eb(x)
            ga = torch.randn(oa.shape, generator=torch.Generator().manual_seed(6)).to(dev).to(oa.dtype)
            gb = torch.randn(ob.shape, generator=torch.Generator().manual_seed(7)).to(dev).to(ob.dtype)
            torch.autograd.backward([oa, ob], [ga, gb])
            res.append((oa.detach(), ob.detach(), ea.embeddings.grad.clone(), eb.embeddings.grad.clone()))
        assert torch.equal(res[0][0], res[1][0]) and torch.equal(res[0][1], res[1][1])
        for k in (2, 3):
            a, b = res[0][k], res[1][k]
            assert a.dtype == torch.float32
            assert float((a - b).abs().max()) <= 1e-5 * float(b.abs().max())


@pytest.mark.parametrize('frozen', [0, 1])
def test_dual_backward_with_a_frozen_table(cuda_lib, dev, frozen):
    """A table with requires_grad False (the stylization stage optimises the colour table only) takes no share of the dual
    scatter: its .grad stays None, the other table's gradient is unchanged."""
    from nerfstyle_b200.gridencoder import grid_encode_dual
    ea, eb = _default_encoder(dev, 1), _default_encoder(dev, 2)
    x = _points(7001, 3, dev, -1.0, 1.0)
    grads = []
    for freeze in (False, True):
        for e in (ea, eb):
            e.embeddings.grad = None
            e.embeddings.requires_grad_(True)
        if freeze:
            (ea, eb)[frozen].embeddings.requires_grad_(False)
        with torch.autocast('cuda', dtype=torch.float16):
            oa, ob = grid_encode_dual(x, ea, eb)
        ga = torch.randn(oa.shape, generator=torch.Generator().manual_seed(6)).to(dev).to(oa.dtype)
        gb = torch.randn(ob.shape, generator=torch.Generator().manual_seed(7)).to(dev).to(ob.dtype)
        outs, gs = ([oa, ob], [ga, gb]) if not freeze else ([(oa, ob)[1 - frozen]], [(ga, gb)[1 - frozen]])
        torch.autograd.backward(outs, gs)
        grads.append([None if e.embeddings.grad is None else e.embeddings.grad.clone() for e in (ea, eb)])
    assert grads[1][frozen] is None and grads[0][frozen] is not None
    a, b = grads[1][1 - frozen], grads[0][1 - frozen]
    assert float((a - b).abs().max()) <= 1e-5 * float(b.abs().max()) and float(b.abs().max()) > 0
    for e in (ea, eb):
        e.embeddings.requires_grad_(True)
    # the C ABI insists that a frozen table drops BOTH its pointers, and that at least one table is live
    assert cuda_lib.nrf_grid_encode_backward_dual(None, 8, 8, 8, 8, 8, 4, 16, 0.5, 16, 0, 1, 0, 1, 0, None, None) == -1
    assert cuda_lib.nrf_grid_encode_backward_dual(None, None, 8, 8, None, None, 4, 16, 0.5, 16, 0, 1, 0, 1, 0, None, None) == -1


def test_dual_input_transform_is_bit_exact(cuda_lib, dev):
    """xform folds BBox.normalize and the encoder's input map into the kernel: same f32 operations -> same bits."""
    from nerfstyle_b200.gridencoder import grid_encode_dual
    ea, eb = _default_encoder(dev, 1), _default_encoder(dev, 2)
    x = _points(9001, 3, dev, -2.2, 2.2)                       # some points outside the box
    bmin = torch.tensor([-2., -2., -2.], device=dev)
    bsize = torch.tensor([4., 4., 4.], device=dev)
    xf = torch.cat([bmin, bsize, torch.ones(1, device=dev)])
    a0, a1 = grid_encode_dual(x, ea, eb, xform=xf)
    b0, b1 = grid_encode_dual((x - bmin) / bsize, ea, eb)
    assert torch.equal(a0, b0) and torch.equal(a1, b1)
    bsize2 = torch.tensor([3.7, 4.1, 4.9], device=dev)          # non-power-of-two sizes: true division per axis
    xf2 = torch.cat([bmin, bsize2, torch.ones(1, device=dev)])
    c0, _ = grid_encode_dual(x, ea, eb, xform=xf2)
    d0, _ = grid_encode_dual((x - bmin) / bsize2, ea, eb)
    assert torch.equal(c0, d0)


def test_backward_skips_exact_zero_gradients(cuda_lib, oracle, dev):
    """Rows whose gradient is exactly zero (samples behind an early-terminated ray) are skipped by the scatter; the
    table gradient is unchanged: whole warps of zeros, zeros inside aggregated runs, and zeros in one channel only."""
    enc = _default_encoder(dev)
    B = 6016
    t = torch.linspace(0, 1, B, device=dev)[:, None]
    x = torch.tensor([[-0.9, -0.7, 0.3]], device=dev) + t * torch.tensor([[1.7, 1.1, 0.4]], device=dev)
    out = enc(x)
    grad = torch.randn(out.shape, generator=torch.Generator().manual_seed(1)).to(dev)
    grad[1000:3000] = 0            # whole warps
    grad[3001:3500:3] = 0          # scattered rows inside runs
    grad[4000:4500, 0::2] = 0      # one channel of every level
    out.backward(grad)
    ege = oracle.grid_encode_backward(grad.cpu().numpy(), ((x + 1) / 2).cpu().numpy(), enc.offsets.cpu().numpy(),
                                      enc.embeddings.shape[0], 2, enc.per_level_scale, 16, 0, True, 0)
    ge = enc.embeddings.grad.cpu().numpy()
    assert np.abs(ge - ege).max() <= 1e-5 * np.abs(ege).max()


@pytest.mark.parametrize('B', [1, 15, 17, 63, 65, 1031, 40007])
@pytest.mark.parametrize('half', [True, False])
def test_pair_backward_walk(cuda_lib, oracle, dev, B, half):
    """nrf_grid_encode_backward_pair's walk form (one lane walks a chunk of consecutive samples at one level and keeps the
    current cell in registers; cells are flushed directly or parked in a shared-memory ring and drained four at a time)
    against the CPU oracle's scatter and the thread-per-sample form, for every chunk length and both flush forms:
    ragged chunk tails, ray-like runs, cell changes at ray boundaries, out-of-range rows, exact-zero gradient rows and
    the in-kernel input transform."""
    enc = _default_encoder(dev)
    g = torch.Generator().manual_seed(B)
    nray = B // 40 + 1
    o = (torch.rand(nray, 3, generator=g) * 3.6 - 1.8).repeat_interleave(40, dim=0)[:B]
    d = torch.nn.functional.normalize(torch.randn(nray, 3, generator=g), dim=-1).repeat_interleave(40, dim=0)[:B]
    x = (o + d * (torch.arange(B)[:, None] % 40) * 3.4e-3).to(dev)             # world coordinates, box [-2, 2]^3
    if B > 20:
        x[7] = 5.0                                                               # outside the box: contributes nothing
    xf = torch.tensor([-2., -2., -2., 4., 4., 4., 1.], device=dev)
    dt = torch.float16 if half else torch.float32
    g0 = torch.randn(B, 32, generator=g).to(dev).to(dt)
    g1 = torch.randn(B, 32, generator=g).to(dev).to(dt)
    if B > 1000:
        g0[100:400] = 0; g1[100:400] = 0                                         # dead samples behind a terminated ray
        g0[500:600:3] = 0
        g1[700:800, 4:6] = 0
    S = float(np.float32(np.log2(enc.per_level_scale)))
    st = torch.cuda.current_stream().cuda_stream
    T = enc.embeddings.shape[0]
    pts = ((x - xf[:3]) / xf[3:6] + 1) * (1 / (2 * xf[6]))                      # what xform computes, operation for operation
    exp = [oracle.grid_encode_backward(gg.float().cpu().numpy(), pts.cpu().numpy(), enc.offsets.cpu().numpy(), T, 2,
                                       enc.per_level_scale, 16, 0, True, 0) for gg in (g0, g1)]
    res = {}
    try:
        for walk, queue in ((0, 1), (16, 0), (32, 0), (64, 0), (32, 1), (64, 1), (128, 1), (256, 1)):
            cuda_lib.nrf_grid_set_bwd_walk(walk)
            cuda_lib.nrf_grid_set_bwd_walk_queue(queue)
            gp = torch.zeros(T, 2, 2, device=dev)
            assert cuda_lib.nrf_grid_encode_backward_pair(g0.data_ptr(), g1.data_ptr(), x.data_ptr(), enc.offsets.data_ptr(), gp.data_ptr(),
                                                          B, 16, S, 16, 0, 1, 0, 1 if half else 0, xf.data_ptr(), st) == 0
            torch.cuda.synchronize()
            res[(walk, queue)] = gp
            for e in (0, 1):
                got = gp[:, e].cpu().numpy()
                assert np.abs(got - exp[e]).max() <= 1e-5 * max(np.abs(exp[e]).max(), 1e-30), (walk, queue, e)
    finally:
        cuda_lib.nrf_grid_set_bwd_walk(128)
        cuda_lib.nrf_grid_set_bwd_walk_queue(1)
    assert float((res[(64, 1)] - res[(0, 1)]).abs().max()) <= 1e-5 * float(res[(0, 1)].abs().max())


@pytest.mark.parametrize('amp', [False, True])
def test_single_table_backward_walk(cuda_lib, oracle, dev, amp):
    """The reference-facing GridEncoder (one table, gridencoder/grid.py:71-97) takes the walk form of the scatter too when
    its preconditions hold: against the CPU oracle and against the thread-per-sample kernel, ray-ordered samples with
    exact-zero gradient rows, ragged length."""
    enc = _default_encoder(dev)
    B = 30011
    g = torch.Generator().manual_seed(4)
    nray = B // 50 + 1
    o = (torch.rand(nray, 3, generator=g) * 1.6 - 0.8).repeat_interleave(50, dim=0)[:B]
    d = torch.nn.functional.normalize(torch.randn(nray, 3, generator=g), dim=-1).repeat_interleave(50, dim=0)[:B]
    x = (o + d * (torch.arange(B)[:, None] % 50) * 1.7e-3).to(dev)
    x[9] = 2.5
    grads = {}
    for walk in (128, 0):
        cuda_lib.nrf_grid_set_bwd_walk(walk)
        try:
            enc.embeddings.grad = None
            with torch.autocast('cuda', dtype=torch.float16, enabled=amp):
                out = enc(x)
            grad = torch.randn(out.shape, generator=torch.Generator().manual_seed(5)).to(dev).to(out.dtype)
            grad[2000:2600] = 0
            grad[4000:4100:2] = 0
            out.backward(grad)
            grads[walk] = enc.embeddings.grad.clone()
        finally:
            cuda_lib.nrf_grid_set_bwd_walk(128)
    assert grads[128].dtype == torch.float32
    ege = oracle.grid_encode_backward(grad.float().cpu().numpy(), ((x + 1) / 2).cpu().numpy(), enc.offsets.cpu().numpy(),
                                      enc.embeddings.shape[0], 2, enc.per_level_scale, 16, 0, True, 0)
    for walk in (128, 0):
        assert np.abs(grads[walk].cpu().numpy() - ege).max() <= 1e-5 * np.abs(ege).max(), walk
