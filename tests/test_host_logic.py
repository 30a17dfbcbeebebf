"""Host-side mirror of the reference interface (no GPU): module surface, parameter names, layouts."""
import os
import sys

import numpy as np
import pytest
import torch

REF = '/root/reference'


def test_gridencoder_layout():
    from nerfstyle_b200.model import get_grid_encoder
    enc = get_grid_encoder(max_bound=4.0)
    assert enc.offsets.dtype == torch.int32 and enc.offsets[-1].item() == 6299960
    assert tuple(enc.embeddings.shape) == (6299960, 2)
    assert enc.n_output_dims == 32 and enc.output_dim == 32
    assert abs(enc.per_level_scale - 2 ** (8 / 15)) < 1e-12
    assert float(enc.embeddings.abs().max()) <= 1e-4
    assert set(enc.state_dict().keys()) == {'embeddings', 'offsets'}


def test_tcnn_network_params():
    from nerfstyle_b200 import tcnn
    cfg = lambda h, a: {'otype': 'FullyFusedMLP', 'activation': 'ReLU', 'output_activation': a, 'n_neurons': 64,  # noqa
                        'n_hidden_layers': h}
    sizes = [tcnn.Network(32, 1, cfg(1, 'None'), 80000).params.numel(), tcnn.Network(32, 8, cfg(1, 'None')).params.numel(),
             tcnn.Network(32, 16, cfg(1, 'None')).params.numel(), tcnn.Network(16, 3, cfg(2, 'Sigmoid')).params.numel()]
    assert sizes == [3072, 3072, 3072, 6144]
    n = tcnn.Network(32, 1, cfg(1, 'None'), 80000)
    assert n.params.dtype == torch.float32 and n.dtype == torch.float16 and n.loss_scale == 128.0
    assert n.n_input_dims == 32 and n.n_output_dims == 1
    a = tcnn.Network(32, 1, cfg(1, 'None'), 1).params
    b = tcnn.Network(32, 1, cfg(1, 'None'), 1).params
    assert torch.equal(a, b)
    with pytest.raises(RuntimeError):
        tcnn.Network(32, 1, dict(cfg(1, 'None'), n_neurons=128))


def test_model_parameter_names():
    from nerfstyle_b200.model import StyleTCNerf
    m = StyleTCNerf([-2., -2., -2.], [2., 2., 2.], class_dim=8)
    names = {n for n, _ in m.named_parameters()}
    assert names == {'x_density_embedder.embeddings', 'x_color_embedder.embeddings', 'density_net.params',
                     'color1_net.params', 'color2_net.params', 'class_net.params'}
    # trainers/base.py:186-198 selects by these substrings
    for kw in ('x_density_embedder', 'x_color_embedder', 'net'):
        assert any(kw in n for n in names)


def test_ops_raise_without_cuda():
    """No CPU fallback: CPU tensors are moved with .cuda() like the reference, which fails loudly without a GPU."""
    if torch.cuda.is_available():
        pytest.skip('has a GPU')
    from nerfstyle_b200 import raymarching
    with pytest.raises((RuntimeError, AssertionError)):
        raymarching.near_far_from_aabb(torch.zeros(4, 3), torch.ones(4, 3), torch.tensor([-1., -1, -1, 1, 1, 1]))


def test_dropin_surface():
    import nerfstyle_b200.dropin as dropin
    rm, ge, tc = dropin.install(force=True)
    import raymarching
    from gridencoder import GridEncoder  # noqa: F401
    import tinycudann as tcnn
    for name in ['near_far_from_aabb', 'sph_from_ray', 'morton3D', 'morton3D_invert', 'packbits', 'march_rays_train',
                 'composite_rays_train', 'march_rays', 'composite_rays']:
        assert callable(getattr(raymarching, name))
    assert hasattr(tcnn, 'Network') and hasattr(tcnn, 'Encoding')
    saved = sys.modules.get('nerf_lib')
    try:
        dropin.install(force=True, nerf_lib=True)
        from nerf_lib import nerf_lib               # renderer.py:9
        assert callable(nerf_lib.generate_rays) and nerf_lib.device is None
    finally:
        sys.modules.pop('nerf_lib', None)
        if saved is not None:
            sys.modules['nerf_lib'] = saved


def test_reference_modules_import_unchanged_on_dropin():
    """The reference's own renderer.py, networks/tcnn_nerf.py, networks/style_nerf.py, common.py, config.py, nerf_lib.py
    and utils/ import and construct against the drop-in (tests/refenv.py: only absent THIRD-PARTY packages -- matplotlib,
    torch_ema, dacite, simple_parsing, imageio -- are stood in for).  Runs wherever the reference is mounted or staged
    (oracle/_ref/pysrc, written by build()); the GPU tests in test_reference_callers_gpu.py then execute these modules."""
    import refenv
    assert refenv.reference_root() is not None, 'run __graft_entry__.build() where /root/reference is mounted'
    with refenv.ReferenceEnv() as E:
        tn, sn = E.tcnn_nerf, E.style_nerf
        assert tn.GridEncoder.__module__ == 'nerfstyle_b200.gridencoder'
        assert E.renderer.raymarching.__nerfstyle_b200__ and E.style_nerf.tcnn.__nerfstyle_b200__
        cfg = E.network_config()
        bbox = E.common.BBox.from_radius(2.0)
        model = sn.StyleTCNerf(cfg, bbox, 8, torch.float16, use_dir=False)
        assert model.x_density_embedder.offsets[-1].item() == 6299960
        assert model.color2_net.params.numel() == 6144
        # the parameter names trainers/base.py:186-198 selects by substring
        assert [n for n, _ in model.named_parameters()] == [
            'x_density_embedder.embeddings', 'x_color_embedder.embeddings', 'density_net.params', 'color1_net.params',
            'color2_net.params', 'class_net.params']
        # use_dir=True builds tcnn.Encoding(SphericalHarmonics, degree 4) and a 32-wide colour-2 input (style_nerf.py:33-42,72-85)
        model_d = sn.StyleTCNerf(cfg, bbox, 8, torch.float16, use_dir=True)
        assert model_d.d_embedder.n_output_dims == 16 and model_d.color2_net.n_input_dims == 32
        # the single-grid TCNerf of networks/tcnn_nerf.py: 16-wide density head, 31-wide rgb input (tcnn_nerf.py:97-122)
        tc = tn.TCNerf(cfg, bbox, torch.float16)
        assert tc.d_embedder.n_output_dims == 16 and tc.density_net.n_output_dims == 16 and tc.rgb_net.n_input_dims == 31
        # the reference's Renderer constructs around it and its state_dict has the reference's keys (renderer.py:78-91)
        intr = E.common.Intrinsics(378, 504, 383.83, 383.83, 252., 189.)
        r = E.renderer.Renderer(model, E.renderer_config(), intr, 2.0, raymarch_channels=11)
        assert r.cascade == 2 and r.density_bitfield.numel() == 524288
        assert sorted(r.state_dict().keys()) == ['bound', 'density_bitfield', 'density_grid', 'intr', 'local_step', 'mean_count',
                                                 'mean_density', 'model', 'precrop_frac', 'raymarch_channels', 'step_counter']
    assert 'renderer' not in sys.modules or not getattr(sys.modules['renderer'], '__file__', '').startswith(refenv.STAGED)


def test_fused_optimizer_pairing_rules_on_cpu(cuda_lib):
    """FusedAdamEMA's host logic without a GPU: which tensors are paired, how the interleaved buffers are laid out, the
    gradient-sink protocol (grad_pair_buffer / grad_of / zero_grad) and the shard plan."""
    from nerfstyle_b200.optim import FusedAdamEMA
    torch.manual_seed(0)
    T = 1 << 19                                      # 2 T elements >= half_copy_min_numel
    a, b = torch.nn.Parameter(torch.randn(T, 2)), torch.nn.Parameter(torch.randn(T, 2))
    w = torch.nn.Parameter(torch.randn(3072))
    opt = FusedAdamEMA([a, w, b], enable_amp=True)
    assert opt.pair_idx == (0, 2)
    assert opt.half_pair.shape == (T, 2, 2) and opt.half_pair.dtype == torch.float16 and opt.half_pair.is_contiguous()
    assert torch.equal(a._nrf_half_copy, a.detach().half()) and torch.equal(b._nrf_half_copy, b.detach().half())
    assert a._nrf_half_pair[0] is b._nrf_half_pair[0] and (a._nrf_half_pair[1], b._nrf_half_pair[1]) == (0, 1)
    assert a._nrf_half_copy.data_ptr() == opt.half_pair.data_ptr() and b._nrf_half_copy.data_ptr() == opt.half_pair.data_ptr() + 4
    assert w._nrf_half_copy.is_contiguous()           # small tensors keep their own fp16 copy
    # writing the parameters by other means + refresh keeps the interleaved copy current
    with torch.no_grad():
        b.mul_(0.5)
    opt.refresh_half_copies()
    assert torch.equal(opt.half_pair[:, 1], b.detach().half())
    # gradient sink: the buffer is created zeroed, stays across backwards of one step, is cleared after zero_grad()
    assert opt.grad_of(a) is None
    gp = opt.grad_pair_buffer()
    assert gp.shape == (T, 2, 2) and gp.dtype == torch.float32 and float(gp.abs().max()) == 0.0 and opt.grad_pair_valid
    gp[:, 1] += 1.0
    assert opt.grad_pair_buffer() is gp and float(gp[:, 1].min()) == 1.0          # accumulation, no re-zeroing
    assert torch.equal(opt.grad_of(b), gp[:, 1]) and float(opt.grad_of(a).abs().max()) == 0.0
    opt.zero_grad()
    assert opt.grad_of(a) is None and not opt.grad_pair_valid
    assert float(opt.grad_pair_buffer().abs().max()) == 0.0
    # no pairing without AMP, with pair_tables off, or when the shapes differ / three tables share a shape
    assert FusedAdamEMA([a, b], enable_amp=False).pair_idx is None
    assert FusedAdamEMA([a, b], pair_tables=False).pair_idx is None
    c = torch.nn.Parameter(torch.randn(T + 8, 2))
    assert FusedAdamEMA([a, c]).pair_idx is None
    assert FusedAdamEMA([a, b, torch.nn.Parameter(torch.randn(T, 2))]).pair_idx is None
    # shard plan: both tables of a pair get the same (even) element range; state lives only for the shard
    sh = FusedAdamEMA([a, w, b], world_size=4, rank=3)
    assert sh.shard[0] == sh.shard[2] == (3 * (2 * T // 4), 2 * T) and sh.shard[1] is None and sh.pair_idx == (0, 2)
    assert sh.exp_avg[0].numel() == 2 * T // 4 and sh.exp_avg[1].numel() == 3072


def test_fused_optimizer_ownership_and_freshness_on_cpu(cuda_lib):
    """the fp16 shadows / pair buffer / gradient sink must not outlive their optimizer, and a shadow must never be
    read stale after the parameter was written (checkpoint load with the optimizer already built)."""
    from nerfstyle_b200 import optim
    from nerfstyle_b200.optim import FusedAdamEMA
    torch.manual_seed(1)
    T = 1 << 19
    a, b = torch.nn.Parameter(torch.randn(T, 2)), torch.nn.Parameter(torch.randn(T, 2))
    w = torch.nn.Parameter(torch.randn(3072))
    o1 = FusedAdamEMA([a, w, b])
    assert optim.optimizer_of(a) is o1 and optim.live_grad_sink(a)[0] is o1
    # a version-visible write (what load_state_dict does) is picked up by the next reader, pair buffer included
    with torch.no_grad():
        a.copy_(torch.full_like(a, 0.25))
        w.mul_(2.0)
    assert a._version != a._nrf_half_version
    h = optim.current_half_copy(a)
    assert float(h.min()) == float(h.max()) == 0.25 and a._version == a._nrf_half_version
    assert torch.equal(o1.half_pair[:, 0], a.detach().half()) and torch.equal(o1.half_pair[:, 1], b.detach().half())
    assert torch.equal(optim.current_half_copy(w), w.detach().half())
    # a write through .data is invisible to the version counter: explicit refresh (ema_scope does this itself)
    a.data.fill_(0.5)
    assert float(optim.current_half_copy(a).max()) == 0.25
    o1.refresh_half_copies()
    assert float(optim.current_half_copy(a).min()) == 0.5
    # ema_scope swaps the EMA weights in (fp16 shadows included) and the training weights back
    o1.ema[0].fill_(-1.0)
    with o1.ema_scope():
        assert float(a.max()) == -1.0 and float(optim.current_half_copy(a).max()) == -1.0
    assert float(a.min()) == 0.5 and float(optim.current_half_copy(a).min()) == 0.5
    # a second optimizer over the same parameters takes them over; the first one refuses to run
    o2 = FusedAdamEMA([a, w, b], pair_tables=False)
    assert not o1.alive and optim.optimizer_of(a) is o2
    assert not hasattr(a, '_nrf_grad_sink') and not hasattr(a, '_nrf_half_pair') and optim.live_grad_sink(a) is None
    with pytest.raises(RuntimeError):
        o1.grad_pair_buffer()
    with pytest.raises(RuntimeError):
        o1.step()
    o2.detach()
    for p in (a, b, w):
        assert not any(hasattr(p, k) for k in optim._ATTRS)
    assert optim.current_half_copy(a) is None
    # the shard plan and the pairing decision depend on (shape, world) only -- identical on every rank
    plans = [FusedAdamEMA([a, w, b], world_size=8, rank=r) for r in (0, 3, 7)]
    assert len({(p.pair_idx, tuple(s is not None for s in p.shard)) for p in plans}) == 1
    odd = [torch.nn.Parameter(torch.randn((1 << 19) + 4, 2)) for _ in range(2)]          # numel / 16 is odd -> not sharded anywhere
    odd_plans = [FusedAdamEMA(odd, world_size=16, rank=r) for r in (0, 1)]
    assert all(p.shard == [None, None] and p.pair_idx == (0, 1) for p in odd_plans)
    # per-parameter learning rates (the reference's second parameter group)
    o3 = FusedAdamEMA([a, w, b], lr=[0.01, 0.005, 0.01])
    assert o3.lrs == [0.01, 0.005, 0.01] and o3.pair_idx == (0, 2)
    assert FusedAdamEMA([a, w, b], lr=[0.01, 0.005, 0.02]).pair_idx is None         # paired tables share one fused pass


def test_pipeline_chunks_cover_all_rows():
    """gridencoder._pipeline_chunks: the row ranges of the gather -> networks pipeline partition [0, B), start on multiples of
    128 rows (the networks' tile size) and collapse to one range for small batches."""
    from nerfstyle_b200.gridencoder import _pipeline_chunks
    assert _pipeline_chunks(4_139_648) == [(0, 4_139_648)]                  # the pipeline is off by default (measured: no gain)
    for B in (0, 1, 127, 1 << 19, (1 << 20) - 1, 1 << 20, 4_139_648, 4_430_363, 33_000_001):
        ch = _pipeline_chunks(B, n=4)
        assert ch[0][0] == 0 and ch[-1][1] == B
        assert all(a[1] == b[0] for a, b in zip(ch[:-1], ch[1:]))
        assert all(r0 % 128 == 0 for r0, _ in ch)
        assert len(ch) == (1 if B < (1 << 20) else min(4, B >> 19))
