"""GPU parity of the ray-marching operator set against the CPU oracle (through the drop-in Python surface, which
calls the C ABI).  Integers / IEEE-only float paths: bit-exact.  Compositing (ex2.approx): rel 1e-4."""
import os

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(__file__), 'golden')


def _rays(n, seed, device):
    from nerfstyle_b200.scenes import random_rays
    return random_rays(n, seed, device)


def _bitfield(kind, device, H=128, C=2):
    from nerfstyle_b200 import raymarching, scenes
    grid = scenes.analytic_density_grid(C, H, 2.0) if kind == 'analytic' else scenes.bernoulli_density_grid(C, H, 0.5, 1)
    return raymarching.packbits(grid.to(device), 0.5), grid


def test_near_far_bit_exact(cuda_lib, oracle, dev):
    from nerfstyle_b200 import raymarching
    o, d = _rays(5000, 0, dev)
    o = o * 6.0                                   # many origins outside the box -> misses
    d[0] = torch.tensor([1.0, 0.0, 0.0]); d[1] = torch.tensor([0.0, -1.0, 0.0]); d[2] = torch.tensor([0.0, 0.0, 1.0])
    aabb = torch.tensor([-2., -2, -2, 2, 2, 2], device=dev)
    nears, fars = raymarching.near_far_from_aabb(o, d, aabb, 0.2)
    en, ef = oracle.near_far_from_aabb(o.cpu().numpy(), d.cpu().numpy(), aabb.cpu().numpy(), 0.2)
    assert np.array_equal(nears.cpu().numpy().view(np.uint32), en.view(np.uint32))
    assert np.array_equal(fars.cpu().numpy().view(np.uint32), ef.view(np.uint32))
    assert (en == np.finfo(np.float32).max).sum() > 100


def test_morton_packbits_bit_exact(cuda_lib, oracle, dev):
    from nerfstyle_b200 import raymarching
    g = torch.Generator().manual_seed(0)
    coords = torch.randint(0, 128, (100003, 3), generator=g, dtype=torch.int32)
    idx = raymarching.morton3D(coords.to(dev))
    assert np.array_equal(idx.cpu().numpy(), oracle.morton3D(coords.numpy()))
    back = raymarching.morton3D_invert(idx)
    assert torch.equal(back.cpu(), coords)
    assert raymarching.morton3D(torch.zeros(0, 3, dtype=torch.int32, device=dev)).numel() == 0
    for shape in [(2, 128 ** 3), (1, 8 * 1027), (3, 8 * 5)]:       # multiple-of-4 and ragged byte counts
        grid = torch.rand(shape, generator=g)
        grid[0, :9] = torch.tensor([0.5, 0.50001, 0.49999, 0.5, 1., 0., 0.5, 0.5, 0.7])
        bits = raymarching.packbits(grid.to(dev), 0.5)
        assert np.array_equal(bits.cpu().numpy(), oracle.packbits(grid.numpy(), 0.5))
    # in-place form
    buf = torch.zeros(2 * 128 ** 3 // 8, dtype=torch.uint8, device=dev)
    grid = torch.rand(2, 128 ** 3, generator=g)
    out = raymarching.packbits(grid.to(dev), 0.3, buf)
    assert out.data_ptr() == buf.data_ptr() and np.array_equal(buf.cpu().numpy(), oracle.packbits(grid.numpy(), 0.3))


@pytest.mark.parametrize('kind', ['analytic', 'bernoulli'])
@pytest.mark.parametrize('max_steps,dt_gamma', [(1024, 0.0), (96, 0.0), (1024, 1.0 / 128)])
def test_march_rays_train_bit_exact(cuda_lib, oracle, dev, kind, max_steps, dt_gamma):
    from nerfstyle_b200 import raymarching
    N = 4096
    o, d = _rays(N, 0, dev)
    bits, _ = _bitfield(kind, dev)
    aabb = torch.tensor([-2., -2, -2, 2, 2, 2], device=dev)
    nears, fars = raymarching.near_far_from_aabb(o, d, aabb, 0.2)
    counter = torch.zeros(2, dtype=torch.int32, device=dev)
    xyzs, dirs, deltas, rays = raymarching.march_rays_train(o, d, None, 2.0, bits, 2, 128, nears, fars, counter, -1, True,
                                                            128, True, dt_gamma, max_steps, False)
    ec = np.zeros(2, np.int32)
    ex, ed, el, er = oracle.march_rays_train(o.cpu().numpy(), d.cpu().numpy(), None, 2.0, bits.cpu().numpy(), 2, 128,
                                             nears.cpu().numpy(), fars.cpu().numpy(), ec, -1, True, 128, True, dt_gamma,
                                             max_steps, False)
    assert np.array_equal(rays.cpu().numpy(), er)                       # ids, offsets, counts
    assert np.array_equal(counter.cpu().numpy(), ec)
    assert xyzs.shape == ex.shape and xyzs.shape[0] % 128 == 0 and xyzs.shape[0] > ec[0]
    for a, b in ((xyzs, ex), (dirs, ed), (deltas, el)):
        assert np.array_equal(a.cpu().numpy().view(np.uint32), b.view(np.uint32))


@pytest.mark.parametrize('mode', [0, 1])
def test_march_modes_and_ndc(cuda_lib, oracle, dev, mode):
    """thread-per-ray (0) and warp-per-ray (1) walkers give identical bits; NDC deltas (thread walker) match the oracle."""
    from nerfstyle_b200 import raymarching
    cuda_lib.nrf_march_set_mode(mode)
    try:
        N = 3001
        o, d = _rays(N, 5, dev)
        o = o * 3.0                                    # some origins outside the box
        bits, _ = _bitfield('bernoulli', dev)
        aabb = torch.tensor([-2., -2, -2, 2, 2, 2], device=dev)
        nears, fars = raymarching.near_far_from_aabb(o, d, aabb, 0.05)
        for is_ndc in (False, True):
            z_hats = (torch.rand(N, generator=torch.Generator().manual_seed(1)) + 0.5).to(dev) if is_ndc else None
            counter = torch.zeros(2, dtype=torch.int32, device=dev)
            xyzs, dirs, deltas, rays = raymarching.march_rays_train(o, d, z_hats, 2.0, bits, 2, 128, nears, fars, counter, -1,
                                                                    False, 128, True, 1.0 / 256, 1024, is_ndc)
            ec = np.zeros(2, np.int32)
            ex, ed, el, er = oracle.march_rays_train(o.cpu().numpy(), d.cpu().numpy(),
                                                     z_hats.cpu().numpy() if is_ndc else None, 2.0, bits.cpu().numpy(), 2, 128,
                                                     nears.cpu().numpy(), fars.cpu().numpy(), ec, -1, False, 128, True, 1.0 / 256,
                                                     1024, is_ndc)
            assert np.array_equal(rays.cpu().numpy(), er) and np.array_equal(counter.cpu().numpy(), ec)
            assert np.array_equal(xyzs.cpu().numpy().view(np.uint32), ex.view(np.uint32))
            a, b = deltas.cpu().numpy(), el
            assert np.array_equal(a[:, :2].view(np.uint32), b[:, :2].view(np.uint32))
            if is_ndc:
                np.testing.assert_allclose(a[:, 2:], b[:, 2:], rtol=1e-5, atol=1e-6)
    finally:
        cuda_lib.nrf_march_set_mode(1)


def test_march_rays_train_edge_cases(cuda_lib, oracle, dev):
    from nerfstyle_b200 import raymarching
    bits, _ = _bitfield('analytic', dev)
    aabb = torch.tensor([-2., -2, -2, 2, 2, 2], device=dev)
    # all rays miss the box
    o = torch.full((257, 3), 10.0, device=dev)
    d = torch.tensor([[1.0, 0.0, 0.0]], device=dev).repeat(257, 1)
    nears, fars = raymarching.near_far_from_aabb(o, d, aabb, 0.2)
    counter = torch.zeros(2, dtype=torch.int32, device=dev)
    xyzs, dirs, deltas, rays = raymarching.march_rays_train(o, d, None, 2.0, bits, 2, 128, nears, fars, counter, -1, False,
                                                            128, True, 0., 1024, False)
    assert counter.tolist() == [0, 257] and xyzs.shape[0] == 128 and float(xyzs.abs().sum()) == 0.0
    assert rays[:, 2].sum().item() == 0
    # mean_count path: fixed-size outputs, rays beyond the capacity dropped (offset + count >= M)
    o, d = _rays(512, 3, dev)
    nears, fars = raymarching.near_far_from_aabb(o, d, aabb, 0.2)
    counter = torch.zeros(2, dtype=torch.int32, device=dev)
    xyzs, dirs, deltas, rays = raymarching.march_rays_train(o, d, None, 2.0, bits, 2, 128, nears, fars, counter, 20000, False,
                                                            128, False, 0., 1024, False)
    ec = np.zeros(2, np.int32)
    ex, ed, el, er = oracle.march_rays_train(o.cpu().numpy(), d.cpu().numpy(), None, 2.0, bits.cpu().numpy(), 2, 128,
                                             nears.cpu().numpy(), fars.cpu().numpy(), ec, 20000, False, 128, False, 0., 1024,
                                             False)
    assert xyzs.shape[0] == 20096 and ex.shape[0] == 20096 and ec[0] > 20096
    assert np.array_equal(rays.cpu().numpy(), er) and np.array_equal(counter.cpu().numpy(), ec)
    assert np.array_equal(xyzs.cpu().numpy().view(np.uint32), ex.view(np.uint32))
    assert np.array_equal(deltas.cpu().numpy().view(np.uint32), el.view(np.uint32))
    # empty batch
    e = torch.zeros(0, 3, device=dev)
    xyzs, dirs, deltas, rays = raymarching.march_rays_train(e, e, None, 2.0, bits, 2, 128, e[:, 0], e[:, 0], None, -1, False,
                                                            128, True, 0., 1024, False)
    assert rays.shape == (0, 3) and xyzs.shape[0] == 0           # zeros(0,3)[:128] in the reference


def test_march_rays_train_hits_max_steps(cuda_lib, oracle, dev):
    """Fully occupied grid + rays along the box diagonals: every ray is cut at max_steps samples."""
    from nerfstyle_b200 import raymarching
    bits = torch.full((2 * 128 ** 3 // 8,), 255, dtype=torch.uint8, device=dev)
    g = torch.Generator().manual_seed(0)
    sgn = (torch.randint(0, 2, (300, 3), generator=g) * 2 - 1).float()
    o = (sgn * 2.5 + torch.randn(300, 3, generator=g) * 0.01).to(dev)
    d = (-sgn / 3 ** 0.5).to(dev)
    aabb = torch.tensor([-2., -2, -2, 2, 2, 2], device=dev)
    nears, fars = raymarching.near_far_from_aabb(o, d, aabb, 0.2)
    counter = torch.zeros(2, dtype=torch.int32, device=dev)
    xyzs, dirs, deltas, rays = raymarching.march_rays_train(o, d, None, 2.0, bits, 2, 128, nears, fars, counter, -1, False, 128,
                                                            True, 0., 64, False)
    ec = np.zeros(2, np.int32)
    ex, ed, el, er = oracle.march_rays_train(o.cpu().numpy(), d.cpu().numpy(), None, 2.0, bits.cpu().numpy(), 2, 128,
                                             nears.cpu().numpy(), fars.cpu().numpy(), ec, -1, False, 128, True, 0., 64, False)
    assert (er[:, 2] == 64).all() and np.array_equal(rays.cpu().numpy(), er)
    assert np.array_equal(xyzs.cpu().numpy().view(np.uint32), ex.view(np.uint32))
    # N*max_steps samples exactly: the last ray trips the reference's `offset + count >= M` drop rule (:517)
    assert ec[0] == 300 * 64 and xyzs.shape[0] == 300 * 64 and float(xyzs[-64:].abs().sum()) == 0.0


def test_golden_fixture_on_gpu(cuda_lib, dev):
    from nerfstyle_b200 import raymarching
    g = np.load(os.path.join(GOLD, 'march_composite.npz'))
    bound, H, C, ms = float(g['bound']), int(g['H']), int(g['C']), int(g['max_steps'])
    t = lambda k: torch.from_numpy(g[k]).to(dev)          # noqa: E731
    aabb = torch.tensor([-bound] * 3 + [bound] * 3, device=dev)
    nears, fars = raymarching.near_far_from_aabb(t('rays_o'), t('rays_d'), aabb, 0.2)
    assert torch.equal(nears.cpu(), torch.from_numpy(g['nears']))
    counter = torch.zeros(2, dtype=torch.int32, device=dev)
    xyzs, dirs, deltas, rays = raymarching.march_rays_train(t('rays_o'), t('rays_d'), None, bound, t('bitfield'), C, H, nears,
                                                            fars, counter, -1, False, 128, True, 0., ms, False)
    assert np.array_equal(rays.cpu().numpy(), g['rays']) and np.array_equal(counter.cpu().numpy(), g['counter'])
    assert np.array_equal(xyzs[:512].cpu().numpy(), g['xyzs']) and np.array_equal(deltas[:512].cpu().numpy(), g['deltas'])
    ws, depth, image = raymarching.composite_rays_train(t('sigmas'), t('rgbs'), deltas, rays, 1e-4, False)
    np.testing.assert_allclose(ws.cpu().numpy(), g['weights_sum'], rtol=1e-4, atol=1e-6)
    np.testing.assert_allclose(image.cpu().numpy(), g['image'], rtol=1e-4, atol=1e-6)
    np.testing.assert_allclose(depth.cpu().numpy(), g['depth'], rtol=1e-4, atol=1e-6)


def _composite_case(dev, C, seed, sigma_scale):
    from nerfstyle_b200 import raymarching
    N = 2048
    o, d = _rays(N, seed, dev)
    bits, _ = _bitfield('analytic', dev)
    aabb = torch.tensor([-2., -2, -2, 2, 2, 2], device=dev)
    nears, fars = raymarching.near_far_from_aabb(o, d, aabb, 0.2)
    counter = torch.zeros(2, dtype=torch.int32, device=dev)
    xyzs, dirs, deltas, rays = raymarching.march_rays_train(o, d, None, 2.0, bits, 2, 128, nears, fars, counter, -1, False, 128,
                                                            True, 0., 1024, False)
    g = torch.Generator().manual_seed(seed)
    M = xyzs.shape[0]
    sigmas = (torch.rand(M, generator=g) * sigma_scale).to(dev)
    rgbs = torch.rand(M, C, generator=g).to(dev)
    return sigmas, rgbs, deltas, rays


@pytest.mark.parametrize('C,sigma_scale', [(11, 2.0), (3, 40.0), (16, 0.1), (40, 5.0)])
def test_composite_rays_train_fwd_bwd(cuda_lib, oracle, dev, C, sigma_scale):
    """rel 1e-4 per ray / sample (north_star tolerance); sigma_scale=40 exercises early termination (T < 1e-4)."""
    from nerfstyle_b200 import raymarching
    sigmas, rgbs, deltas, rays = _composite_case(dev, C, 1, sigma_scale)
    sigmas.requires_grad_(True)
    rgbs.requires_grad_(True)
    ws, depth, image = raymarching.composite_rays_train(sigmas, rgbs, deltas, rays, 1e-4, False)
    g = torch.Generator().manual_seed(5)
    gws = torch.randn(ws.shape, generator=g).to(dev)
    gim = torch.randn(image.shape, generator=g).to(dev)
    (ws * gws).sum().add((image * gim).sum()).add(depth.sum() * 3.0).backward()   # grad_depth must be ignored
    n = lambda x: x.detach().cpu().numpy()      # noqa: E731
    ews, edepth, eimage = oracle.composite_rays_train_forward(n(sigmas), n(rgbs), n(deltas), n(rays), 1e-4, False)
    np.testing.assert_allclose(n(ws), ews, rtol=1e-4, atol=2e-6)
    np.testing.assert_allclose(n(image), eimage, rtol=1e-4, atol=2e-6)
    np.testing.assert_allclose(n(depth), edepth, rtol=1e-4, atol=2e-5)
    egs, egr = oracle.composite_rays_train_backward(n(gws), n(gim), n(sigmas), n(rgbs), n(deltas), n(rays), ews, eimage, 1e-4)
    gs, gr = n(sigmas.grad), n(rgbs.grad)
    # a sample exactly at the T threshold may terminate one step apart on the two ex2 implementations:
    # allow a handful of rays to differ, everything else within tolerance
    bad_r = ~np.isclose(gr, egr, rtol=1e-4, atol=2e-6).all(axis=1)
    bad_s = ~np.isclose(gs, egs, rtol=2e-3, atol=2e-5)
    assert bad_r.sum() <= 4 and bad_s.sum() <= 4, (bad_r.sum(), bad_s.sum())
    scale = np.abs(egs).max()
    assert np.abs(gs - egs)[~bad_s].max() <= 1e-4 * scale + 2e-5


@pytest.mark.parametrize('T_thresh', [0.0, 1e-2])
def test_inference_loop_matches_oracle(cuda_lib, oracle, dev, T_thresh):
    """renderer.py:237-293 loop.  T_thresh = 0: nothing depends on the GPU's ex2.approx, so the alive sets are identical
    step by step (exact) and the accumulators agree to rel 1e-4.  T_thresh = 1e-2: a ray whose transmittance sits on the
    threshold may die one iteration apart on the two exp implementations; the images must still agree to 2 * T_thresh."""
    from nerfstyle_b200 import raymarching
    N, C = 3000, 5
    o, d = _rays(N, 2, dev)
    bits, _ = _bitfield('analytic', dev)
    aabb = torch.tensor([-2., -2, -2, 2, 2, 2], device=dev)
    nears, fars = raymarching.near_far_from_aabb(o, d, aabb, 0.2)
    n = lambda x: x.detach().cpu().numpy().copy()      # noqa: E731
    o_n, d_n, bits_n, nears_n, fars_n = n(o), n(d), n(bits), n(nears), n(fars)

    def field_fn(xyzs, dirs):
        sig = xyzs.norm(dim=-1) * 3.0 + 0.5
        rgb = torch.cat([dirs.abs(), xyzs[:, :2].abs()], dim=-1)
        return sig, rgb

    # GPU loop
    ws = torch.zeros(N, device=dev); depth = torch.zeros(N, device=dev); image = torch.zeros(N, C, device=dev)
    alive = torch.arange(N, dtype=torch.int32, device=dev)
    rays_t = nears.clone()[:, None]
    gpu_alive_hist, step = [], 0
    while step < 1024 and len(alive) > 0:
        n_alive = len(alive)
        n_step = max(min(N // n_alive, 8), 1)
        xyzs, dirs, deltas = raymarching.march_rays(n_alive, n_step, alive, rays_t, o, d, None, 2.0, bits, 2, 128, nears, fars,
                                                    128, False, 0., 1024, False)
        assert xyzs.shape[0] % 128 == 0 and xyzs.shape[0] > n_alive * n_step - 1
        sig, rgb = field_fn(xyzs, dirs)
        raymarching.composite_rays(n_alive, n_step, alive, rays_t, sig, rgb, deltas, False, ws, depth, image, T_thresh)
        alive2, k = raymarching.compact_rays_alive(alive)
        alive = alive[alive >= 0]
        assert torch.equal(alive2, alive) and k == len(alive)
        gpu_alive_hist.append(n(alive))
        step += n_step
    # oracle loop
    ews, edepth, eimage = np.zeros(N, np.float32), np.zeros(N, np.float32), np.zeros((N, C), np.float32)
    ealive = np.arange(N, dtype=np.int32)
    erays_t = nears_n.copy()[:, None]
    cpu_alive_hist, step = [], 0
    while step < 1024 and len(ealive) > 0:
        n_alive = len(ealive)
        n_step = max(min(N // n_alive, 8), 1)
        ex, ed, el = oracle.march_rays(n_alive, n_step, ealive, erays_t, o_n, d_n, None, 2.0, bits_n, 2, 128, nears_n, fars_n, 128,
                                       False, 0., 1024, False)
        sig, rgb = field_fn(torch.from_numpy(ex), torch.from_numpy(ed))
        oracle.composite_rays(n_alive, n_step, ealive, erays_t, sig.numpy(), rgb.numpy(), el, False, ews, edepth, eimage, T_thresh)
        ealive = np.ascontiguousarray(ealive[ealive >= 0])
        cpu_alive_hist.append(ealive.copy())
        step += n_step
    assert len(gpu_alive_hist) > 10
    if T_thresh == 0.0:
        assert len(gpu_alive_hist) == len(cpu_alive_hist)
        for a, b in zip(gpu_alive_hist, cpu_alive_hist):
            assert np.array_equal(a, b)
        np.testing.assert_allclose(n(ws), ews, rtol=1e-4, atol=2e-6)
        np.testing.assert_allclose(n(image), eimage, rtol=1e-4, atol=2e-6)
        np.testing.assert_allclose(n(depth), edepth, rtol=1e-4, atol=2e-5)
        np.testing.assert_allclose(n(rays_t), erays_t, rtol=0, atol=0)
    else:
        close = np.isclose(n(image), eimage, rtol=1e-4, atol=2e-6).all(axis=1)
        assert (~close).mean() < 0.01                      # a handful of threshold-straddling rays
        assert np.abs(n(image) - eimage).max() <= 2 * T_thresh * max(1.0, float(np.abs(eimage).max()))
        assert np.abs(n(ws) - ews).max() <= 2 * T_thresh


def test_plain_c_host_of_the_abi(cuda_lib, dev):
    """examples/c_host.c drives the C ABI with nothing but the CUDA runtime (no Python objects, no torch types): built
    with gcc against include/nerfstyle_b200.h and run as a separate process."""
    import os
    import shutil
    import subprocess
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    exe = os.path.join(root, 'examples', 'c_host')
    src = os.path.join(root, 'examples', 'c_host.c')
    if shutil.which('gcc') is None:
        pytest.skip('no gcc')
    if not os.path.exists(exe) or os.path.getmtime(exe) < os.path.getmtime(src):
        subprocess.check_call(['gcc', '-O2', '-I', os.path.join(root, 'include'), '-I', '/usr/local/cuda/include', src, '-o', exe,
                               '-L', os.path.join(root, 'nerfstyle_b200'), '-lnerfstyle_b200', '-L', '/usr/local/cuda/lib64', '-lcudart',
                               '-Wl,-rpath,' + os.path.join(root, 'nerfstyle_b200'), '-lm'])
    out = subprocess.run([exe], capture_output=True, text=True, timeout=120)
    assert out.returncode == 0 and 'C HOST OK' in out.stdout, out.stdout + out.stderr


def test_composite_backward_without_zero_fill_equals_reference_path(cuda_lib, dev):
    """Rays produced by this package's march_rays_train partition the sample rows exactly, so composite_rays_train's backward
    skips the reference's two zero fills (raymarching.py:339-340) and lets the kernel write the zeros of terminated / dropped
    samples itself.  Same gradients, bit for bit, with early termination active and the allocator's free blocks poisoned."""
    from nerfstyle_b200 import raymarching
    o, d = _rays(3000, 4, dev)
    bits, _ = _bitfield('analytic', dev)
    aabb = torch.tensor([-2., -2, -2, 2, 2, 2], device=dev)
    nears, fars = raymarching.near_far_from_aabb(o, d, aabb, 0.2)
    counter = torch.zeros(2, dtype=torch.int32, device=dev)
    xyzs, dirs, deltas, rays = raymarching.march_rays_train(o, d, None, 2.0, bits, 2, 128, nears, fars, counter, -1, False, 128, True,
                                                            0., 1024, False)
    assert getattr(rays, '_nrf_dense_rows', None) == int(counter[0]) and xyzs.shape[0] > int(counter[0])
    M = xyzs.shape[0]
    g = torch.Generator(device=dev).manual_seed(9)
    sig0 = torch.rand(M, device=dev, generator=g) * 40          # dense medium: most rays terminate early
    rgb0 = torch.rand(M, 11, device=dev, generator=g)
    gw, gi = torch.randn(3000, device=dev, generator=g), torch.randn(3000, 11, device=dev, generator=g)
    res = []
    for dense in (True, False):
        r = rays if dense else rays.clone()                      # the clone carries no tag -> zero-filled reference path
        sig, rgb = sig0.clone().requires_grad_(True), rgb0.clone().requires_grad_(True)
        ws, depth, image = raymarching.composite_rays_train(sig, rgb, deltas, r, 1e-4, False)
        poison = torch.full((M * 12 + 4096,), float('nan'), device=dev)
        del poison                                               # freed NaN blocks are what torch.empty will hand back
        torch.autograd.backward([ws, image], [gw, gi])
        res.append((sig.grad.clone(), rgb.grad.clone()))
    assert torch.equal(res[0][0], res[1][0]) and torch.equal(res[0][1], res[1][1])
    assert bool(torch.isfinite(res[0][1]).all()) and float((res[0][0] == 0).float().mean()) > 0.3    # many terminated samples
