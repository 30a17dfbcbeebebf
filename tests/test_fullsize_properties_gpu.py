"""Size-independent properties of the hot path at BASELINE.json's FULL sizes (configs 2 and 3: 8192 room-shaped rays with
~4 M samples, the 1008 x 756 frame with 762 048 rays), where the CPU oracle is too slow to be the checker: partition /
sortedness of the compaction offsets, occupancy of every emitted sample, determinism (idempotence), linearity of the
compositing and of the hash-grid encoder in their linear arguments, adjointness of the encoder's forward and backward,
Morton / packbits round trips over the whole grid, agreement of the two MLP arithmetic modes, and tile-sharding invariance
of the inference loop (what the multi-GPU render relies on)."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

BOUND, H, CAS = 2.0, 128, 2


def _room_batch(dev, n_rays=8192, seed=0):
    from nerfstyle_b200 import scenes
    intr = dict(scenes.ROOM)
    pose = scenes.synthetic_poses(intr['n_train'], 0)[3]
    idx = scenes.frame_indices(intr, n_rays, torch.Generator().manual_seed(seed)).to(dev)
    return scenes.generate_rays(pose, intr, dev, idx)


def _field_occupancy(dev):
    """The bitfield a random-init field gets from update_state at step 0 (about half the cells)."""
    from nerfstyle_b200 import model as M
    torch.manual_seed(0)
    m = M.StyleTCNerf([-BOUND] * 3, [BOUND] * 3, class_dim=8).to(dev)
    r = M.Renderer(m, BOUND, raymarch_channels=11).to(dev)
    with torch.autocast('cuda', dtype=torch.float16):
        r.update_state()
    return m, r


def test_march_rays_train_full_batch_properties(cuda_lib, dev):
    from nerfstyle_b200 import raymarching
    m, r = _field_occupancy(dev)
    o, d = _room_batch(dev)
    N = o.shape[0]
    nears, fars = raymarching.near_far_from_aabb(o, d, r.aabb, 0.2)
    outs = []
    for rep in range(2):
        counter = torch.zeros(2, dtype=torch.int32, device=dev)
        outs.append(raymarching.march_rays_train(o, d, None, BOUND, r.density_bitfield, CAS, H, nears, fars, counter, -1, True, 128, True,
                                                 0., 1024, False) + (counter,))
    (xyzs, dirs, deltas, rays, counter) = outs[0]
    for a, b in zip(outs[0], outs[1]):
        assert torch.equal(a, b)                                              # idempotent / deterministic
    total = int(counter[0])
    assert total > 2_000_000 and int(counter[1]) == N
    cnt, off = rays[:, 2].long(), rays[:, 1].long()
    assert torch.equal(rays[:, 0].long(), torch.arange(N, device=dev))         # ray ids in order
    assert torch.equal(off, torch.cumsum(cnt, 0) - cnt)                       # offsets = exclusive scan: a partition of [0, total)
    assert int(cnt.sum()) == total and int(cnt.max()) <= 1024
    m_pad = total + (128 - total % 128)
    assert xyzs.shape[0] == m_pad and float(xyzs[total:].abs().max()) == 0.0 and float(deltas[total:].abs().max()) == 0.0
    x = xyzs[:total]
    assert float(x.abs().max()) <= BOUND
    dt_min = np.float32(2 * np.sqrt(3.0)) / np.float32(1024)
    assert bool((deltas[:total, 0] == float(dt_min)).all())                   # dt_gamma = 0: every step is dt_min
    assert bool((deltas[:total, 1] >= float(dt_min) * 0.999).all())           # distance to the previous sample >= one step
    # every emitted sample sits in an occupied cell of the cascade its position selects (raymarching.cu:460-479)
    mx = x.abs().amax(dim=1)
    level = (mx >= 1.0).long()                                                # C = 2: frexp exponent clamped to [0, 1]
    rb = torch.where(level == 1, torch.full_like(mx, 0.5), torch.ones_like(mx))
    cell = (0.5 * (x * rb[:, None] + 1.0) * H).clamp(0, H - 1).to(torch.int32)
    idx = raymarching.morton3D(cell).long() + level * H ** 3
    bits = (r.density_bitfield[idx >> 3].long() >> (idx & 7)) & 1
    assert bool((bits == 1).all())
    # a ray's samples advance monotonically along its direction
    ray_of = torch.repeat_interleave(torch.arange(N, device=dev), cnt)
    t = ((x - o[ray_of]) * d[ray_of]).sum(dim=1)
    same = ray_of[1:] == ray_of[:-1]
    assert bool((t[1:][same] > t[:-1][same] - 1e-4).all())


def test_composite_linearity_and_bounds_at_full_size(cuda_lib, dev):
    from nerfstyle_b200 import raymarching
    m, r = _field_occupancy(dev)
    o, d = _room_batch(dev, seed=1)
    nears, fars = raymarching.near_far_from_aabb(o, d, r.aabb, 0.2)
    counter = torch.zeros(2, dtype=torch.int32, device=dev)
    xyzs, dirs, deltas, rays = raymarching.march_rays_train(o, d, None, BOUND, r.density_bitfield, CAS, H, nears, fars, counter, -1, True,
                                                            128, True, 0., 1024, False)
    M_ = xyzs.shape[0]
    g = torch.Generator().manual_seed(3)
    sig = (torch.rand(M_, generator=g) * 4).to(dev)
    a = torch.rand(M_, 11, generator=g).to(dev)
    b = torch.rand(M_, 11, generator=g).to(dev)
    ws_a, dep_a, img_a = raymarching.composite_rays_train(sig, a, deltas, rays, 1e-4, False)
    ws_b, dep_b, img_b = raymarching.composite_rays_train(sig, b, deltas, rays, 1e-4, False)
    ws_c, dep_c, img_c = raymarching.composite_rays_train(sig, 0.25 * a + 3.0 * b, deltas, rays, 1e-4, False)
    assert torch.equal(ws_a, ws_b) and torch.equal(dep_a, dep_c)              # weights / depth do not depend on the colours
    assert float((img_c - (0.25 * img_a + 3.0 * img_b)).abs().max()) <= 2e-5 * float(img_c.abs().max())
    assert float(ws_a.max()) <= 1.0 + 1e-5 and float(ws_a.min()) >= 0.0 and float(dep_a.min()) >= 0.0
    ones = torch.ones(M_, 11, device=dev)
    _, _, img_1 = raymarching.composite_rays_train(sig, ones, deltas, rays, 1e-4, False)
    assert float((img_1 - ws_a[:, None]).abs().max()) <= 1e-5                 # compositing a constant 1 gives the weight sum
    # backward: grad_rgbs of <image, G> is w x G, so summed over a ray's samples and contracted with G it is bounded by ws |G|^2
    a_ = a.clone().requires_grad_(True)
    s_ = sig.clone().requires_grad_(True)
    ws, dep, img = raymarching.composite_rays_train(s_, a_, deltas, rays, 1e-4, False)
    G = torch.rand(img.shape, generator=g).to(dev)
    (img * G).sum().backward()
    assert torch.isfinite(a_.grad).all() and torch.isfinite(s_.grad).all()
    assert float(a_.grad.min()) >= 0.0                                        # w >= 0, G >= 0
    total = int(counter[0])
    assert float(a_.grad[total:].abs().max()) == 0.0 and float(s_.grad[total:].abs().max()) == 0.0      # padding rows


def test_grid_encoder_linearity_and_adjointness_at_full_size(cuda_lib, dev):
    """fp32 tables, 4 M points: the encoder is linear in its table -- enc(x; A + B) = enc(x; A) + enc(x; B) -- and its
    backward is the adjoint of that linear map: <enc(x; T), G> = <T, scatter(x, G)>."""
    from nerfstyle_b200.gridencoder import GridEncoder, grid_encode
    torch.manual_seed(1)
    enc = GridEncoder(num_levels=16, level_dim=2, per_level_scale=np.exp2(np.log2(4096 / 16) / 15), base_resolution=16, log2_hashmap_size=19,
                      align_corners=True).to(dev)
    B_ = 4_000_000
    x = torch.rand(B_, 3, device=dev) * 0.5 + 0.5                              # the [0.5, 1] octant the model addresses
    A = torch.randn_like(enc.embeddings)
    Bt = torch.randn_like(enc.embeddings)

    def run(table):
        return grid_encode(x, table, enc.offsets, enc.per_level_scale, enc.base_resolution, False, 0, True, 0)
    with torch.no_grad():
        ea, eb, eab = run(A), run(Bt), run(A + Bt)
    assert float((eab - (ea + eb)).abs().max()) <= 1e-5 * float(eab.abs().max())
    T = A.clone().requires_grad_(True)
    y = run(T)
    G = torch.randn(y.shape, device=dev)
    lhs = float((y.detach().double() * G.double()).sum())
    y.backward(G)
    rhs = float((T.detach().double() * T.grad.double()).sum())
    assert abs(lhs - rhs) <= 1e-5 * abs(lhs), (lhs, rhs)
    # rows of levels the points cannot reach in the [0.5, 1] octant receive exactly zero gradient; touched rows are finite
    assert torch.isfinite(T.grad).all()
    touched = (T.grad != 0).any(dim=1)
    assert 0 < int(touched.sum()) < T.shape[0]


@pytest.mark.parametrize('half_grads', [True, False])
def test_paired_walk_scatter_is_the_adjoint_of_the_paired_gather_at_full_size(cuda_lib, dev, half_grads):
    """4.2 M ray-ordered samples (8192 rays x 512 march steps: the cell runs the walk kernels aggregate): the paired
    scatter -- ring form for f16 gradient rows, register-flush form for f32 rows -- is the adjoint of the paired f32 gather,
    <enc(x; T0, T1), (G0, G1)> = <(T0, T1), grad_pair>, and agrees with the thread-per-sample kernel."""
    from nerfstyle_b200.model import get_grid_encoder
    enc = get_grid_encoder(max_bound=4.0).to(dev)
    g = torch.Generator().manual_seed(11)
    n_rays, n_steps = 8192, 512
    o = (torch.rand(n_rays, 1, 3, generator=g) * 0.3 + 0.5).to(dev)
    d = torch.nn.functional.normalize(torch.rand(n_rays, 1, 3, generator=g) + 0.05, dim=-1).to(dev)
    t = (torch.arange(n_steps, device=dev, dtype=torch.float32) * 4.2e-4)[None, :, None]
    x = (o + d * t).reshape(-1, 3).contiguous()                                  # a few rays leave [0, 1]^3: those rows contribute nothing
    B_ = x.shape[0]
    T_ = enc.embeddings.shape[0]
    pair = torch.randn(T_, 2, 2, generator=g).to(dev)                             # [row][table][2] f32
    S = float(np.float32(np.log2(enc.per_level_scale)))
    st = torch.cuda.current_stream().cuda_stream
    o0, o1 = torch.empty(B_, 32, device=dev), torch.empty(B_, 32, device=dev)
    assert cuda_lib.nrf_grid_encode_forward_pair(x.data_ptr(), pair.data_ptr(), enc.offsets.data_ptr(), o0.data_ptr(), o1.data_ptr(),
                                                 B_, 16, S, 16, 0, 1, 0, 0, None, None, None, st) == 0
    dt = torch.float16 if half_grads else torch.float32
    G0 = torch.randn(B_, 32, device=dev).to(dt)
    G1 = torch.randn(B_, 32, device=dev).to(dt)
    lhs = float((o0.double() * G0.double()).sum() + (o1.double() * G1.double()).sum())
    out = {}
    try:
        for walk in (128, 0):
            cuda_lib.nrf_grid_set_bwd_walk(walk)
            gp = torch.zeros(T_, 2, 2, device=dev)
            assert cuda_lib.nrf_grid_encode_backward_pair(G0.data_ptr(), G1.data_ptr(), x.data_ptr(), enc.offsets.data_ptr(), gp.data_ptr(),
                                                          B_, 16, S, 16, 0, 1, 0, 1 if half_grads else 0, None, st) == 0
            torch.cuda.synchronize()
            out[walk] = gp
    finally:
        cuda_lib.nrf_grid_set_bwd_walk(128)
    rhs = float((pair.double() * out[128].double()).sum())
    assert abs(lhs - rhs) <= 1e-5 * abs(lhs), (lhs, rhs)
    assert float((out[128] - out[0]).abs().max()) <= 2e-5 * float(out[0].abs().max())
    assert torch.isfinite(out[128]).all() and 0 < int((out[128] != 0).any(dim=2).any(dim=1).sum()) < T_


def test_morton_and_packbits_over_the_whole_grid(cuda_lib, dev):
    from nerfstyle_b200 import raymarching
    idx = torch.arange(H ** 3, dtype=torch.int32, device=dev)
    coords = raymarching.morton3D_invert(idx)
    assert int(coords.min()) == 0 and int(coords.max()) == H - 1
    assert torch.equal(raymarching.morton3D(coords), idx)                     # a bijection of [0, 128^3)
    key = coords[:, 0].long() * H * H + coords[:, 1].long() * H + coords[:, 2].long()
    assert int(torch.unique(key).numel()) == H ** 3
    grid = torch.rand(CAS, H ** 3, generator=torch.Generator().manual_seed(2)).to(dev)
    for thresh in (0.0, 0.37, 0.999, 2.0):
        bits = raymarching.packbits(grid, thresh)
        assert bits.numel() == CAS * H ** 3 // 8
        unpacked = ((bits.long()[:, None] >> torch.arange(8, device=dev)) & 1).reshape(-1)
        assert torch.equal(unpacked.bool(), (grid.reshape(-1) > thresh))


def test_mlp_arithmetic_modes_agree_at_full_size(cuda_lib, dev):
    """4 M rows through the tensor-core network (f16 operands) and the fp32 parity mode: the outputs differ only by the f16
    roundings of the perf mode, everywhere."""
    from nerfstyle_b200 import tcnn
    net = tcnn.Network(32, 16, {'otype': 'FullyFusedMLP', 'activation': 'ReLU', 'output_activation': 'None', 'n_neurons': 64,
                                'n_hidden_layers': 1}, seed=5).to(dev)
    x = (torch.randn(4_000_037, 32, generator=torch.Generator().manual_seed(4)) * 0.5).to(dev).half()
    with torch.no_grad():
        y16 = net(x).float()
        with tcnn.parity_mode():
            y32 = net(x)
    scale = float(y32.abs().max())
    assert float((y16 - y32).abs().max()) <= 4e-3 * scale
    assert float((y16 - y32).abs().mean()) <= 3e-4 * scale


def test_full_frame_render_is_deterministic_and_tile_invariant(cuda_lib, dev):
    """The 1008 x 756 frame of config 3 through the device-driven loop: same frame twice -> same bits; rendering the two
    halves of the rows separately (what each rank of a 2-GPU render does) reproduces the full-frame image up to the
    n_step schedule's effect (<= 5e-5, tests/test_pipeline_gpu.py::test_render_test_graph_equals_reference_loop)."""
    from nerfstyle_b200 import model as M, raymarching, scenes
    torch.manual_seed(0)
    m = M.StyleTCNerf([-BOUND] * 3, [BOUND] * 3, class_dim=8).to(dev)
    r = M.Renderer(m, BOUND, raymarch_channels=11, density_scale=50.0).to(dev)
    r.density_bitfield = raymarching.packbits(scenes.analytic_density_grid(CAS, H, BOUND).to(dev), 0.5)
    intr = scenes.scaled_intrinsics(1008, 756)
    pose = scenes.synthetic_poses(2, 1)[1]
    o, d = scenes.generate_rays(pose, intr, dev, torch.arange(1008 * 756, device=dev))
    with torch.no_grad(), torch.autocast('cuda', dtype=torch.float16):
        a = [t.clone() for t in r.render_test_graph(o, d)]
        b = [t.clone() for t in r.render_test_graph(o, d)]
        half = o.shape[0] // 2
        top = [t.clone() for t in r.render_test_graph(o[:half].contiguous(), d[:half].contiguous())]
        bot = [t.clone() for t in r.render_test_graph(o[half:].contiguous(), d[half:].contiguous())]
    for x, y in zip(a, b):
        assert torch.equal(x, y)
    rgb = a[0]
    assert rgb.shape == (762048, 3) and torch.isfinite(rgb).all()
    assert float(rgb.min()) >= -1e-5 and float(rgb.max()) <= 1.0 + 1e-3
    assert float(a[1].min()) >= 0.0
    for full, t, bt in zip(a, top, bot):
        assert float((full - torch.cat([t, bt])).abs().max()) <= 5e-5
