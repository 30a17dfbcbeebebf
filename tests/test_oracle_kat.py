"""Pins the CPU oracle: closed-form known-answer values derived from the reference formulas,
the committed golden fixtures (regression), and domain properties.  CPU only."""
import os

import numpy as np
import pytest

GOLD = os.path.join(os.path.dirname(__file__), 'golden')


def test_morton_kat(oracle):
    # raymarching.cu:56-71
    coords = [[1, 0, 0], [0, 1, 0], [0, 0, 1], [3, 5, 7], [100, 50, 25], [64, 0, 127], [127, 127, 127]]
    expect = [1, 2, 4, 431, 387156, 1460516, 2097151]
    assert oracle.morton3D(coords).tolist() == expect
    assert oracle.morton3D_invert(expect).tolist() == coords


def test_morton_roundtrip_all_cells(oracle):
    idx = np.arange(128 ** 3, dtype=np.int32)
    assert np.array_equal(oracle.morton3D(oracle.morton3D_invert(idx)), idx)


def test_fast_hash_kat(oracle):
    # gridencoder.cu:35-52, primes 1, 2654435761, 805459861, style prime 3674653429
    cases = {(0, 1, 0): 2654435761, (0, 0, 1): 805459861, (1, 1, 1): 2922720805, (16, 16, 16): 3813859920,
             (123, 456, 789): 635907338, (4095, 4096, 4096): 1390563327}
    mods = [489905, 153493, 339493, 189008, 470282, 151551]
    for (p, h), m in zip(cases.items(), mods):
        assert oracle.fast_hash3(*p) == h
        assert h % 524288 == m
    assert oracle.fast_hash3(1, 1, 1, 3) == 1059153658
    assert 1059153658 % 2 ** 19 == 91898


def test_default_grid_layout(oracle):
    # Appendix B of SURVEY.md: bound=2 -> max_res 4096
    pls = np.exp2(np.log2(4096 / 16) / 15)
    offs, _ = oracle.grid_offsets(3, 16, 2, pls, 16, 19, None, True)
    assert offs.tolist()[:7] == [0, 4096, 17920, 57224, 174880, 532792, 1057080]
    assert offs[-1] == 6299960
    S = np.float32(np.log2(pls))
    res = [oracle.level_resolution(l, S, 16) for l in range(16)]
    assert res == [16, 23, 33, 48, 70, 101, 147, 212, 307, 445, 645, 933, 1351, 1955, 2830, 4096]


def test_grid_index_path_kat(oracle):
    # full kernel_grid index path at kernel-space input (0.75, 0.625, 0.9) on the default grid
    pls = np.exp2(np.log2(4096 / 16) / 15)
    offs, _ = oracle.grid_offsets(3, 16, 2, pls, 16, 19, None, True)
    emb = np.zeros((int(offs[-1]), 2), np.float32)
    _, _, idx = oracle.grid_encode_forward(np.array([[0.75, 0.625, 0.9]], np.float32), emb, offs, pls, 16, False, 0, True,
                                           0, return_indices=True)
    assert idx[0, 0].tolist() == [2752, 2753, 177, 176, 349, 348, 2860, 2861]
    assert idx[1, 0].tolist() == [1051, 1048, 10986, 10985, 7302, 7301, 12919, 12916]
    assert idx[5, 0].tolist() == [276646, 276641, 304745, 304750, 250675, 250676, 216572, 216571]
    assert idx[15, 0].tolist() == [126302, 126303, 91375, 91374, 230643, 230642, 200002, 200003]


def test_grid_weights_kat(oracle):
    # level 0: frac = (0, 0, 0.39999962) -> only corners 0 (w=0.6) and 4 (w=0.4) contribute
    pls = np.exp2(np.log2(4096 / 16) / 15)
    offs, _ = oracle.grid_offsets(3, 16, 2, pls, 16, 19, None, True)
    emb = np.zeros((int(offs[-1]), 2), np.float32)
    emb[2752] = [1.0, 10.0]
    emb[349] = [100.0, 1000.0]
    out, _ = oracle.grid_encode_forward(np.array([[0.75, 0.625, 0.9]], np.float32), emb, offs, pls, 16, False, 0, True, 0)
    np.testing.assert_allclose(out[0, :2], [0.6 * 1 + 0.4 * 100, 0.6 * 10 + 0.4 * 1000], rtol=2e-6)


def test_packbits_kat(oracle):
    grid = np.zeros((1, 16), np.float32)
    grid[0, [0, 3, 9]] = [1.0, 0.5, 2.0]          # strict '>' : 0.5 > 0.5 is False
    assert oracle.packbits(grid, 0.5).tolist() == [0b00000001, 0b00000010]


def test_step_constants(oracle):
    # dt_min = 2*sqrt(3)/1024 and the PTX constant 0x405DB3D7
    two_sqrt3 = np.float32(2.0) * np.float32(1.7320508075688772)
    assert two_sqrt3.view(np.uint32) == 0x405DB3D7
    # a ray through a fully occupied grid takes steps of exactly dt_min
    H, C = 16, 1
    bits = np.full(C * H ** 3 // 8, 255, np.uint8)
    o = np.array([[0.0, 0.0, 0.0]], np.float32)
    d = np.array([[1.0, 0.0, 0.0]], np.float32)
    xyzs, dirs, deltas, rays = oracle.march_rays_train(o, d, None, 1.0, bits, C, H, np.array([0.2], np.float32),
                                                       np.array([1.0], np.float32), None, -1, False, -1, True, 0., 1024)
    assert deltas[0, 0] == np.float32(two_sqrt3 / np.float32(1024))
    assert rays[0].tolist()[0:2] == [0, 0]
    n = rays[0, 2]
    assert abs(n - (1.0 - 0.2) / deltas[0, 0]) <= 1
    # FLT_MAX for a miss
    nears, fars = oracle.near_far_from_aabb(np.array([[5., 5., 5.]], np.float32), np.array([[1., 0., 0.]], np.float32),
                                            np.array([-1, -1, -1, 1, 1, 1], np.float32))
    assert nears[0] == np.finfo(np.float32).max and fars[0] == np.finfo(np.float32).max


def test_golden_march_composite(oracle):
    g = np.load(os.path.join(GOLD, 'march_composite.npz'))
    bound, H, C, ms = float(g['bound']), int(g['H']), int(g['C']), int(g['max_steps'])
    aabb = np.array([-bound] * 3 + [bound] * 3, np.float32)
    nears, fars = oracle.near_far_from_aabb(g['rays_o'], g['rays_d'], aabb, 0.2)
    assert np.array_equal(nears, g['nears']) and np.array_equal(fars, g['fars'])
    counter = np.zeros(2, np.int32)
    xyzs, dirs, deltas, rays = oracle.march_rays_train(g['rays_o'], g['rays_d'], None, bound, g['bitfield'], C, H, nears,
                                                       fars, counter, -1, False, 128, True, 0., ms, False)
    assert np.array_equal(rays, g['rays']) and np.array_equal(counter, g['counter'])
    assert xyzs.shape[0] == int(g['n_rows']) and xyzs.shape[0] % 128 == 0 and xyzs.shape[0] > counter[0]
    assert np.array_equal(xyzs[:512], g['xyzs']) and np.array_equal(deltas[:512], g['deltas'])
    # offsets are the exclusive scan of the counts in ray order
    assert np.array_equal(rays[:, 1], np.concatenate([[0], np.cumsum(rays[:-1, 2])]))
    ws, depth, image = oracle.composite_rays_train_forward(g['sigmas'], g['rgbs'], deltas, rays, 1e-4, False)
    np.testing.assert_allclose(ws, g['weights_sum'], rtol=1e-6, atol=1e-7)
    np.testing.assert_allclose(image, g['image'], rtol=1e-6, atol=1e-7)
    gs, gr = oracle.composite_rays_train_backward(g['grad_ws'], g['grad_image'], g['sigmas'], g['rgbs'], deltas, rays, ws,
                                                  image, 1e-4, False)
    np.testing.assert_allclose(gs, g['grad_sigmas'], rtol=1e-5, atol=1e-6)
    np.testing.assert_allclose(gr, g['grad_rgbs'], rtol=1e-6, atol=1e-7)


def test_golden_grid(oracle):
    g = np.load(os.path.join(GOLD, 'grid_small.npz'))
    offs, pls = g['offsets'], float(g['per_level_scale'])
    emb = np.random.RandomState(int(g['emb_seed'])).uniform(-1, 1, (int(offs[-1]), 2)).astype(np.float32)
    out, _, idx = oracle.grid_encode_forward(g['inputs'], emb, offs, pls, 16, False, 0, True, 0, return_indices=True)
    assert np.array_equal(idx, g['indices'])
    assert np.array_equal(out, g['outputs'])
    assert np.all(out[3] == 0)                      # out-of-range input -> zeros
    ge = oracle.grid_encode_backward(g['grad'], g['inputs'], offs, int(offs[-1]), 2, pls, 16, 0, True, 0)
    np.testing.assert_allclose(ge, g['grad_embeddings'], rtol=1e-6, atol=1e-7)


def test_composite_matches_closed_form(oracle):
    """weights telescope: sum_i w_i = 1 - prod(1 - alpha_i) when nothing terminates early."""
    rs = np.random.RandomState(0)
    n = 40
    sig = rs.rand(n).astype(np.float32)
    dl = np.zeros((n + 88, 4), np.float32)
    dl[:n, 0] = 0.01
    dl[:n, 1] = 0.01
    sig_p = np.zeros(n + 88, np.float32)
    sig_p[:n] = sig
    rgb = np.ones((n + 88, 3), np.float32)
    rays = np.array([[0, 0, n]], np.int32)
    ws, depth, image = oracle.composite_rays_train_forward(sig_p, rgb, dl, rays, 1e-4, False)
    alpha = 1 - np.exp(-sig.astype(np.float64) * 0.01)
    np.testing.assert_allclose(ws[0], 1 - np.prod(1 - alpha), rtol=1e-5)
    np.testing.assert_allclose(image[0], ws[0], rtol=1e-6)


def test_composite_backward_is_gradient(oracle):
    """finite differences of the forward reproduce the analytic backward (away from the T threshold)."""
    rs = np.random.RandomState(1)
    n, C = 12, 4
    M = 128
    sig = np.zeros(M, np.float32); sig[:n] = rs.rand(n) * 3
    rgb = np.zeros((M, C), np.float32); rgb[:n] = rs.rand(n, C)
    dl = np.zeros((M, 4), np.float32); dl[:n, 0] = 0.05; dl[:n, 1] = 0.05
    rays = np.array([[0, 0, n]], np.int32)
    gws = rs.randn(1).astype(np.float32); gim = rs.randn(1, C).astype(np.float32)
    ws, depth, image = oracle.composite_rays_train_forward(sig, rgb, dl, rays, 1e-6, False)
    gs, gr = oracle.composite_rays_train_backward(gws, gim, sig, rgb, dl, rays, ws, image, 1e-6, False)

    def f(s, r):
        w, _, im = oracle.composite_rays_train_forward(s, r, dl, rays, 1e-6, False)
        return float(gws[0] * w[0] + (gim[0] * im[0]).sum())
    eps = 1e-2
    for i in (0, 5, n - 1):
        sp, sm = sig.copy(), sig.copy(); sp[i] += eps; sm[i] -= eps
        fd = (f(sp, rgb) - f(sm, rgb)) / (2 * eps)
        assert abs(fd - gs[i]) < 2e-3 * max(1.0, abs(fd)), (i, fd, gs[i])
        rp, rm = rgb.copy(), rgb.copy(); rp[i, 1] += eps; rm[i, 1] -= eps
        fd = (f(sig, rp) - f(sig, rm)) / (2 * eps)
        assert abs(fd - gr[i, 1]) < 2e-3 * max(1.0, abs(fd))


def test_grid_backward_is_adjoint(oracle):
    """<encode(x; T), g> is linear in T, so backward(g) must equal its exact adjoint."""
    rs = np.random.RandomState(5)
    offs, pls = oracle.grid_offsets(3, 4, 2, 2, 8, 10, desired_resolution=64, align_corners=True)
    T1 = rs.randn(int(offs[-1]), 2).astype(np.float32)
    x = rs.rand(50, 3).astype(np.float32)
    g = rs.randn(50, 8).astype(np.float32)
    out, _ = oracle.grid_encode_forward(x, T1, offs, pls, 8, False, 0, True, 0)
    ge = oracle.grid_encode_backward(g, x, offs, int(offs[-1]), 2, pls, 8, 0, True, 0)
    np.testing.assert_allclose((out.astype(np.float64) * g).sum(), (ge.astype(np.float64) * T1).sum(), rtol=1e-5)


def test_inference_loop_matches_train_march(oracle):
    """march_rays/composite_rays iterated to exhaustion visits exactly the samples march_rays_train emits."""
    g = np.load(os.path.join(GOLD, 'march_composite.npz'))
    bound, H, C, ms = float(g['bound']), int(g['H']), int(g['C']), int(g['max_steps'])
    o, d, bits, nears, fars = g['rays_o'], g['rays_d'], g['bitfield'], g['nears'], g['fars']
    N = o.shape[0]
    rays = g['rays']
    ws = np.zeros(N, np.float32); depth = np.zeros(N, np.float32); image = np.zeros((N, 3), np.float32)
    alive = np.arange(N, dtype=np.int32)
    rays_t = nears.copy()[:, None]
    emitted = np.zeros(N, np.int64)
    step = 0
    while step < ms and len(alive) > 0:
        n_alive = len(alive)
        n_step = max(min(N // n_alive, 8), 1)
        xyzs, dirs, deltas = oracle.march_rays(n_alive, n_step, alive, rays_t, o, d, None, bound, bits, C, H, nears, fars,
                                               128, False, 0., ms, False)
        cnt = (deltas[:n_alive * n_step, 0] > 0).reshape(n_alive, n_step).sum(1)
        np.add.at(emitted, alive, cnt)
        sig = np.zeros(xyzs.shape[0], np.float32)       # zero density: nothing terminates early
        rgb = np.zeros((xyzs.shape[0], 3), np.float32)
        oracle.composite_rays(n_alive, n_step, alive, rays_t, sig, rgb, deltas, False, ws, depth, image, 1e-4)
        alive = np.ascontiguousarray(alive[alive >= 0])
        step += n_step
    assert np.array_equal(emitted, rays[:, 2].astype(np.int64))


def test_spherical_harmonics_basis_is_orthonormal():
    """The SH restatement (oracle.field.sh_encode) is pinned by a closed-form property: the 16 basis functions are
    orthonormal over the unit sphere (Gauss-Legendre in cos(theta) x uniform in phi integrates degree-6 polynomials exactly)."""
    import numpy as np
    from oracle import field
    ct, wt = np.polynomial.legendre.leggauss(16)
    phi = (np.arange(32) + 0.5) * (2 * np.pi / 32)
    CT, PH = np.meshgrid(ct, phi, indexing='ij')
    ST = np.sqrt(1 - CT ** 2)
    d = np.stack([ST * np.cos(PH), ST * np.sin(PH), CT], axis=-1).reshape(-1, 3)
    w = (wt[:, None] * np.full((1, 32), 2 * np.pi / 32)).reshape(-1)
    Y = field.sh_encode((d + 1) / 2, 4)
    G = (Y * w[:, None]).T @ Y
    assert np.abs(G - np.eye(16)).max() < 1e-12


def test_occupancy_oracle_known_answers(oracle):
    """oracle/occupancy.py (renderer.py:120-194): closed-form checks of the sample points and the masked decay / max update."""
    from oracle import occupancy as occ
    H, C, bound = 8, 2, 2.0
    consts = occ.cascade_constants(C, bound, H)
    assert consts[0] == (np.float32(1.0 - 1.0 / 8), np.float32(1.0 / 8)) and consts[1] == (np.float32(2.0 - 2.0 / 8), np.float32(2.0 / 8))
    mid = np.full((C, H ** 3, 3), 0.5, np.float32)                   # noise 0.5 -> no jitter: the scaled cell centre
    pts = occ.points_full_morton(mid, C, bound, H)
    # Morton index 0 is cell (0,0,0) -> centre -1; index 7 is cell (1,1,1); the last index is cell (H-1,)*3 -> centre +1
    assert np.array_equal(pts[0, 0], np.full(3, -consts[0][0], np.float32))
    assert np.array_equal(pts[1, H ** 3 - 1], np.full(3, consts[1][0], np.float32))
    c1 = np.float32(2.0) * np.float32(1.0) * (np.float32(1.0) / np.float32(7.0)) - np.float32(1.0)
    assert np.array_equal(pts[0, 7], np.full(3, c1 * consts[0][0], np.float32))
    assert np.array_equal(pts[0, 1], np.array([c1 * consts[0][0], -consts[0][0], -consts[0][0]], np.float32))   # x is the lowest Morton bit
    lo = occ.points_full_morton(np.zeros_like(mid), C, bound, H)
    assert np.allclose(pts - lo, np.array([c[1] for c in consts], np.float32)[:, None, None])                    # jitter spans +- half a cell
    grid = np.array([[0.0, 1.0, -1.0, 2.0]], np.float32)
    tmp = np.array([[0.5, -1.0, 3.0, 1.0]], np.float32)
    g, mean, bits = occ.grid_update(np.tile(grid, (1, 2)), np.tile(tmp, (1, 2)), 0.95, 10.0)
    assert np.array_equal(g[0, :4], np.array([0.5, 1.0, -1.0, np.float32(2.0) * np.float32(0.95)], np.float32))
    assert abs(mean - (0.5 + 1.0 + 0.0 + 1.9) / 4) < 1e-6 and np.asarray(bits).reshape(-1)[0] == 0b10101010
