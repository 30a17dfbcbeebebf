"""End-to-end parity of one train step (BASELINE config 1: 4096 rays, 16-level 2^19 hash grids x2, 4 MLPs,
composite fwd+bwd) through the drop-in modules against the CPU oracle pipeline."""
import math

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def _build(dev, half_tables, n_rays=4096, table_std=0.5, fused_heads=True):
    from nerfstyle_b200 import model as M, raymarching, scenes
    from oracle import field
    of = field.OracleField(bound=2.0, n_classes=8, half=half_tables, seed=0, table_std=table_std)
    m = M.StyleTCNerf([-2., -2., -2.], [2., 2., 2.], class_dim=8, fused_heads=fused_heads).to(dev)
    with torch.no_grad():
        for n, p in m.named_parameters():
            p.copy_(of.params[n].detach().to(dev))
    r = M.Renderer(m, 2.0, raymarch_channels=11).to(dev)
    r.update_occ = False
    grid = scenes.analytic_density_grid(2, 128, 2.0)
    bits = raymarching.packbits(grid.to(dev), 0.5)
    r.density_bitfield = bits
    o, d = scenes.random_rays(n_rays, 0, dev)
    return of, m, r, o, d, bits


@pytest.mark.parametrize('fused_heads', [True, False])
@pytest.mark.parametrize('amp', [False, True])
def test_train_step_matches_oracle(cuda_lib, oracle, dev, amp, fused_heads):
    """fused_heads=False is the reference's exact op sequence on the drop-in modules; True folds trunc_exp / cat / casts
    into the MLP kernels (tcnn.density_head / color_heads).  Both must match the oracle pipeline."""
    from oracle import field
    of, m, r, o, d, bits = _build(dev, half_tables=amp, fused_heads=fused_heads)
    g = torch.Generator().manual_seed(9)
    target = torch.rand(o.shape[0], 3, generator=g)
    tcls = torch.randint(0, 8, (o.shape[0],), generator=g)
    with torch.autocast('cuda', dtype=torch.float16, enabled=amp):
        image, depth, classes = r.render_train(o, d)
        loss = torch.mean((image - target.to(dev)) ** 2) + 0.001 * torch.nn.functional.cross_entropy(classes, tcls.to(dev))
    scale = 65536.0     # GradScaler's initial scale; the reference always trains under GradScaler; the fp16 MLP backward needs the headroom
    (loss * scale).backward()
    out = field.render_train(of, o.cpu().numpy(), d.cpu().numpy(), bits.cpu().numpy(), 2, 128, 2.0)
    eloss = field.train_step_loss(out, target, tcls)
    (eloss * scale).backward()
    # integers: bit-exact
    n_s = int(out['counter'][0])
    assert n_s > 100000
    # floats
    img, eimg = image.detach().float().cpu().numpy(), out['rgb'].detach().numpy()
    mse = float(np.mean((img - eimg) ** 2))
    # tolerances = 3x the errors measured on the B200 (round 2): image 3.0e-6, classes 1.2e-6, loss 1.7e-6 relative,
    # gradients <= 2.3e-4 of each tensor's scale (the oracle models every fp16 rounding point of the perf-mode MLP)
    assert np.abs(img - eimg).max() < 1e-5, np.abs(img - eimg).max()
    psnr_delta = abs(10 * math.log10(max(np.mean((img - target.numpy()) ** 2), 1e-12)) -
                     10 * math.log10(max(np.mean((eimg - target.numpy()) ** 2), 1e-12)))
    assert psnr_delta < 0.01, psnr_delta
    assert abs(float(loss) - float(eloss)) < 1e-5 * abs(float(eloss)), (float(loss), float(eloss), mse)
    np.testing.assert_allclose(classes.detach().float().cpu().numpy(), out['classes'].detach().numpy(), rtol=0, atol=5e-6)
    # gradients of every parameter (normalised max error)
    for name, p in m.named_parameters():
        gp = p.grad.float().cpu().numpy()
        eg = of.params[name].grad.numpy()
        denom = np.abs(eg).max()
        assert denom > 0, name
        tol = 7e-4
        assert np.abs(gp - eg).max() <= tol * denom, (name, np.abs(gp - eg).max() / denom)


@pytest.mark.parametrize('fused_heads', [True, False])
def test_train_step_parity_mode_rel_1e4(cuda_lib, oracle, dev, fused_heads):
    """The whole train step in fp32 -- fp32 hash tables (no autocast) and the fp32 PARITY-MODE MLP (tcnn.parity_mode()) --
    against the all-fp32 oracle pipeline, at the tolerance BASELINE.json's north_star states: image / class outputs and
    EVERY parameter gradient within rel 1e-4 of the tensor's scale, PSNR delta < 0.01 dB."""
    from nerfstyle_b200 import tcnn
    from oracle import field
    of, m, r, o, d, bits = _build(dev, half_tables=False, fused_heads=fused_heads)
    of.mlp_half = False
    g = torch.Generator().manual_seed(9)
    target = torch.rand(o.shape[0], 3, generator=g)
    tcls = torch.randint(0, 8, (o.shape[0],), generator=g)
    with tcnn.parity_mode():
        image, depth, classes = r.render_train(o, d)
        loss = torch.mean((image - target.to(dev)) ** 2) + 0.001 * torch.nn.functional.cross_entropy(classes, tcls.to(dev))
        loss.backward()
    out = field.render_train(of, o.cpu().numpy(), d.cpu().numpy(), bits.cpu().numpy(), 2, 128, 2.0)
    eloss = field.train_step_loss(out, target, tcls)
    eloss.backward()
    img, eimg = image.detach().cpu().numpy(), out['rgb'].detach().numpy()
    assert np.abs(img - eimg).max() <= 1e-4 * np.abs(eimg).max(), np.abs(img - eimg).max()
    ecls = out['classes'].detach().numpy()
    assert np.abs(classes.detach().cpu().numpy() - ecls).max() <= 1e-4 * np.abs(ecls).max()
    np.testing.assert_allclose(depth.detach().cpu().numpy(), out['depth'].detach().numpy(), rtol=0, atol=1e-4)
    assert abs(float(loss) - float(eloss)) <= 1e-5 * abs(float(eloss))
    psnr = lambda a: 10 * math.log10(max(np.mean((a - target.numpy()) ** 2), 1e-12))          # noqa: E731
    assert abs(psnr(img) - psnr(eimg)) < 0.01
    for name, p in m.named_parameters():
        gp, eg = p.grad.float().cpu().numpy(), of.params[name].grad.numpy()
        err = np.abs(gp - eg).max() / np.abs(eg).max()
        assert err <= 1e-4, (name, err)                  # measured <= 1.5e-5


@pytest.mark.parametrize('amp', [False, True])
def test_fused_heads_equal_reference_op_sequence(cuda_lib, dev, amp):
    """The fused heads reproduce the unfused op sequence: same values (the f16 rounding points are kept), gradients
    within the f16 rounding of dy (the fused path does not round the compositing gradient to f16 first)."""
    outs = []
    for fused in (True, False):
        of, m, r, o, d, bits = _build(dev, half_tables=amp, n_rays=1024, fused_heads=fused)
        with torch.autocast('cuda', dtype=torch.float16, enabled=amp):
            image, depth, classes = r.render_train(o, d)
            loss = torch.mean(image ** 2) + 0.01 * torch.mean(classes ** 2)
        (loss * 65536.0).backward()
        outs.append((image.detach(), classes.detach(), {n: p.grad.clone() for n, p in m.named_parameters()}))
    torch.testing.assert_close(outs[0][0], outs[1][0], rtol=1e-5, atol=1e-6)
    torch.testing.assert_close(outs[0][1], outs[1][1], rtol=1e-5, atol=1e-6)
    for n in outs[0][2]:
        a, b = outs[0][2][n].float(), outs[1][2][n].float()
        assert float((a - b).abs().max()) <= 1e-2 * float(b.abs().max()), n


def test_render_test_matches_render_train(cuda_lib, dev):
    """The inference loop (march_rays/composite_rays) and the training path render the same image when no ray
    terminates early (T_thresh = 0 in both)."""
    of, m, r, o, d, bits = _build(dev, half_tables=False, n_rays=1500, table_std=0.2)
    r.t_thresh = 0.0
    with torch.no_grad():
        a_img, a_depth, a_cls = r.render_train(o, d)
        b_img, b_depth, b_cls = r.render_test(o, d)
    torch.testing.assert_close(a_img, b_img, rtol=1e-4, atol=2e-5)
    torch.testing.assert_close(a_cls, b_cls, rtol=1e-3, atol=1e-4)


def test_render_test_sync_light_loop_equals_reference_loop(cuda_lib, dev):
    """sync_every > 1 (device-side compaction with dead slots, host count refreshed every k iterations) renders the same
    image as the reference's loop (count read back every iteration), with early termination active."""
    of, m, r, o, d, bits = _build(dev, half_tables=False, n_rays=3000, table_std=0.5)
    r.density_scale = 20.0
    with torch.no_grad():
        a = r.render_test(o, d, sync_every=1)
        for k in (2, 4, 7):
            b = r.render_test(o, d, sync_every=k)
            for x, y in zip(a, b):
                torch.testing.assert_close(x, y, rtol=1e-6, atol=1e-7)
    # the compaction primitive itself: survivors in order, -1 tail, device-side count
    from nerfstyle_b200 import raymarching
    v = torch.tensor([5, -1, 7, -1, -1, 9, 11], dtype=torch.int32, device=dev)
    out, cnt = raymarching.compact_rays_alive_nosync(v)
    assert out.tolist() == [5, 7, 9, 11, -1, -1, -1] and int(cnt) == 4


def test_update_state_and_steps(cuda_lib, dev):
    """Occupancy update + a few optimiser steps run end to end and reduce the loss."""
    from nerfstyle_b200 import model as M, scenes
    torch.manual_seed(0)
    m = M.StyleTCNerf([-2., -2., -2.], [2., 2., 2.], class_dim=8).to(dev)
    r = M.Renderer(m, 2.0, raymarch_channels=11).to(dev)
    opt = torch.optim.Adam(m.parameters(), lr=1e-2, eps=1e-15)
    intr = dict(scenes.ROOM)
    pose = scenes.synthetic_poses(2, 0)[0]
    gen = torch.Generator().manual_seed(0)
    losses = []
    for it in range(6):
        idx = scenes.frame_indices(intr, 1024, gen).to(dev)
        o, d = scenes.generate_rays(pose, intr, dev, idx)
        tgt, seg = scenes.synthetic_target(idx, intr)
        with torch.autocast('cuda', dtype=torch.float16):
            img, depth, cls = r.render_train(o, d)
            loss = torch.mean((img - tgt) ** 2)
        opt.zero_grad()
        (loss * 128).backward()
        for p in m.parameters():
            p.grad.div_(128)
        opt.step()
        losses.append(float(loss))
    assert r.local_step == 6 and r.mean_density > 0 and int(r.density_bitfield.count_nonzero()) > 0
    assert all(np.isfinite(losses)) and losses[-1] < losses[0]


def test_fused_optimizer_matches_torch(cuda_lib, dev):
    """FusedAdamEMA == GradScaler + torch.optim.Adam(eps=1e-15) + LambdaLR + torch_ema-style EMA, incl. inf-skip."""
    from nerfstyle_b200.optim import FusedAdamEMA
    torch.manual_seed(0)
    shapes = [(2_000_003,), (3072,), (50, 7)]
    pa = [torch.nn.Parameter(torch.randn(s, device=dev) * 0.1) for s in shapes]
    pb = [torch.nn.Parameter(p.detach().clone()) for p in pa]
    fused = FusedAdamEMA(pa, lr=0.01, lr_decay_steps=50, ema_decay=0.95, init_scale=1024.0, growth_interval=3,
                         half_copy_min_numel=1 << 20)
    opt = torch.optim.Adam(pb, lr=0.01, betas=(0.9, 0.999), eps=1e-15)
    sched = torch.optim.lr_scheduler.LambdaLR(opt, lambda it: 0.1 ** (it / 50))
    scaler = torch.amp.GradScaler('cuda', init_scale=1024.0, growth_interval=3)
    ema = [p.detach().clone() for p in pb]
    n_upd = 0
    scaler.scale(torch.ones(1, device=dev))        # lazily initialises the scaler's device state
    for it in range(9):
        grads = [torch.randn_like(p) for p in pa]
        if it == 4:
            grads[1][5] = float('inf')                  # this step must be skipped and the scale halved
        s_a = float(fused.scale.item())
        for p, q, g in zip(pa, pb, grads):
            p.grad = g * s_a
            q.grad = g * scaler.get_scale()
        assert s_a == scaler.get_scale()
        fused.step()
        scaler.step(opt)
        old = scaler.get_scale()
        scaler.update()
        if old <= scaler.get_scale():
            sched.step()
        n_upd += 1
        decay = min(0.95, (1 + n_upd) / (10 + n_upd))
        with torch.no_grad():
            for e, q in zip(ema, pb):
                e.sub_((1 - decay) * (e - q))
    for p, q in zip(pa, pb):
        torch.testing.assert_close(p, q, rtol=2e-5, atol=2e-6)
    for e, f in zip(ema, fused.ema):
        torch.testing.assert_close(f, e, rtol=2e-5, atol=2e-6)
    assert int(fused.good_steps.item()) == 8
    torch.testing.assert_close(pa[0]._nrf_half_copy.float(), pa[0].detach().half().float())
    for p in pa[1:]:       # the small (MLP-sized) tensors keep a current fp16 copy in every mode
        torch.testing.assert_close(p._nrf_half_copy.float(), p.detach().half().float())


def test_trainstep_async_loss_readback(cuda_lib, dev):
    """TrainStep.step(loss_host=pinned): the loss lands in pinned host memory through a side stream right after the
    forward; the value equals the returned device loss."""
    from nerfstyle_b200 import model as M, scenes
    from nerfstyle_b200.trainer import TrainStep
    torch.manual_seed(0)
    m = M.StyleTCNerf([-2., -2., -2.], [2., 2., 2.], class_dim=8).to(dev)
    r = M.Renderer(m, 2.0, raymarch_channels=11).to(dev)
    ts = TrainStep(r, enable_amp=True)
    intr = dict(scenes.ROOM)
    pose = scenes.synthetic_poses(2, 0)[0]
    gen = torch.Generator().manual_seed(0)
    pin = torch.zeros(1).pin_memory()
    for it in range(3):
        idx = scenes.frame_indices(intr, 512, gen).to(dev)
        o, d = scenes.generate_rays(pose, intr, dev, idx)
        tgt, seg = scenes.synthetic_target(idx, intr)
        loss = ts.step(o, d, tgt, seg, loss_host=pin)
        ts.loss_ready.synchronize()
        assert abs(float(pin[0]) - float(loss)) < 1e-7 and np.isfinite(float(pin[0]))


def test_render_test_graph_equals_reference_loop(cuda_lib, dev):
    """The device-driven loop (control block on the device, a pair of iterations replayed from a CUDA graph) renders
    the same image as the reference's host-driven loop -- first frame (eager pair + capture) and later frames (replay
    only), with early termination active.  steps_per_iteration=1 is the reference's n_step schedule: same sample positions,
    same image to f32 rounding.  The default budget (4 N rows per iteration) re-synchronises a ray's marching time with the
    accumulated (rounded) deltas every 4 samples instead of every sample (composite_rays writes rays_t += sum of deltas,
    raymarching.cu:1184-1226), which moves samples by ~1e-7 and the image by <= 2e-5 -- the same effect the reference's own
    n_step growth (up to 8 as rays die) has."""
    of, m, r, o, d, bits = _build(dev, half_tables=True, n_rays=3000, table_std=0.5)
    r.density_scale = 20.0
    with torch.no_grad(), torch.autocast('cuda', dtype=torch.float16):
        a = r.render_test(o, d, sync_every=1)
        for frame in range(3):
            b = r.render_test_graph(o, d, steps_per_iteration=1)
            for x, y in zip(a, b):
                torch.testing.assert_close(x, y, rtol=1e-5, atol=1e-6)
        it1 = int(r._gs['ctl'][6])
        for spi in (4, 8):
            b = r.render_test_graph(o, d, steps_per_iteration=spi)
            for x, y in zip(a, b):
                assert float((x - y).abs().max()) <= 5e-5, (spi, float((x - y).abs().max()))
            assert int(r._gs['ctl'][6]) < it1 / 2                          # far fewer iterations
        o2, d2 = o.flip(0).contiguous(), d.flip(0).contiguous()          # another frame through the same captured graph
        a2 = r.render_test(o2, d2, sync_every=1)
        b2 = r.render_test_graph(o2, d2, steps_per_iteration=1)
        for x, y in zip(a2, b2):
            torch.testing.assert_close(x, y, rtol=1e-5, atol=1e-6)
    assert int(r._gs['ctl'][0]) == 0 and int(r._gs['ctl'][6]) > 0


def test_paired_tables_match_unpaired(cuda_lib, dev):
    """FusedAdamEMA(pair_tables=True): the two hash tables share one interleaved fp16 copy and one interleaved f32 gradient
    buffer (nrf_grid_encode_forward_pair / _backward_pair, strided nrf_adam_step_ex).  Forward values are identical to
    the unpaired path, gradients equal up to the order of the float atomics, and the optimizer step lands on the same
    parameters / EMA / fp16 copies."""
    from nerfstyle_b200 import model as M, scenes
    from nerfstyle_b200.trainer import TrainStep
    intr = dict(scenes.ROOM)
    pose = scenes.synthetic_poses(2, 0)[0]
    runs = []
    for pair in (True, False):
        torch.manual_seed(0)
        m = M.StyleTCNerf([-2., -2., -2.], [2., 2., 2.], class_dim=8).to(dev)
        with torch.no_grad():                          # tables large enough for non-trivial outputs
            for e in (m.x_density_embedder, m.x_color_embedder):
                e.embeddings.uniform_(-0.5, 0.5, generator=torch.Generator(device=dev).manual_seed(5))
        r = M.Renderer(m, 2.0, raymarch_channels=11).to(dev)
        ts = TrainStep(r, enable_amp=True, pair_tables=pair)
        f = ts.fused
        assert (f.pair_idx is not None) == pair
        if pair:
            assert m.x_density_embedder.embeddings._nrf_half_pair[0] is m.x_color_embedder.embeddings._nrf_half_pair[0]
            torch.testing.assert_close(m.x_color_embedder.embeddings._nrf_half_copy.float(),
                                       m.x_color_embedder.embeddings.detach().half().float(), rtol=0, atol=0)
        gen = torch.Generator().manual_seed(0)
        idx = scenes.frame_indices(intr, 2048, gen).to(dev)
        o, d = scenes.generate_rays(pose, intr, dev, idx)
        tgt, seg = scenes.synthetic_target(idx, intr)
        # one manual forward / backward through the same path TrainStep.step takes
        with torch.autocast('cuda', dtype=torch.float16):
            image, depth, classes = r.render_train(o, d)
            loss, _ = ts.loss_fn(image, classes, tgt, seg)
        f.zero_grad()
        f.scale_loss(loss).backward()
        grads = {n: f.grad_of(p).detach().clone() for n, p in m.named_parameters()}
        if pair:
            assert m.x_density_embedder.embeddings.grad is None and f.grad_pair_valid
        before = {n: p.detach().clone() for n, p in m.named_parameters()}
        f.step()
        runs.append(dict(image=image.detach().clone(), loss=float(loss), grads=grads, before=before,
                         after={n: p.detach().clone() for n, p in m.named_parameters()},
                         half={n: p._nrf_half_copy.float().clone() for n, p in m.named_parameters()},
                         ema=[e.clone() for e in f.ema]))
        # two more ordinary steps must stay finite and keep the fp16 copies current
        for it in range(2):
            idx = scenes.frame_indices(intr, 2048, gen).to(dev)
            o, d = scenes.generate_rays(pose, intr, dev, idx)
            tgt, seg = scenes.synthetic_target(idx, intr)
            assert np.isfinite(float(ts.step(o, d, tgt, seg)))
        for n, p in m.named_parameters():
            torch.testing.assert_close(p._nrf_half_copy.float(), p.detach().half().float(), rtol=0, atol=0)
        assert int(f.good_steps.item()) == 3
    a, b = runs
    assert torch.equal(a['image'], b['image']) and a['loss'] == b['loss']          # same gathers, same arithmetic
    for n in a['grads']:
        ga, gb = a['grads'][n], b['grads'][n]
        assert float((ga - gb).abs().max()) <= 1e-5 * float(gb.abs().max()) + 1e-30, n
    # the step itself: where the two runs' gradients agree in sign and are not rounding noise, Adam's first step is
    # -lr * sign(g) in both; compare on the rows that received a meaningful gradient
    for n in a['after']:
        g = b['grads'][n]
        mask = g.abs() > 1e-3 * g.abs().max()
        da, db = (a['after'][n] - a['before'][n])[mask], (b['after'][n] - b['before'][n])[mask]
        torch.testing.assert_close(da, db, rtol=1e-3, atol=1e-6)
        assert mask.sum() > 0


def _occ_renderers(dev, **kw):
    """Two renderers sharing one model (so both query the same field): fused and reference occupancy update."""
    from nerfstyle_b200 import model as M
    torch.manual_seed(0)
    m = M.StyleTCNerf([-2., -2., -2.], [2., 2., 2.], class_dim=8).to(dev)
    with torch.no_grad():
        for e in (m.x_density_embedder, m.x_color_embedder):
            e.embeddings.uniform_(-0.5, 0.5, generator=torch.Generator(device=dev).manual_seed(5))
    a = M.Renderer(m, 2.0, raymarch_channels=11, fused_occupancy=True, **kw).to(dev)
    b = M.Renderer(m, 2.0, raymarch_channels=11, fused_occupancy=False, **kw).to(dev)
    return m, a, b


def test_update_state_fused_full_phase_equals_reference_ops(cuda_lib, dev, monkeypatch):
    """Renderer.update_state as device passes (nrf_occ_points_full / nrf_occ_update / nrf_packbits_dev) against the reference's
    op sequence (renderer.py:138-194) on the same jitter: the density grid is bit-identical, the mean within 1e-6, the bitfield
    identical up to cells that sit on the threshold."""
    from nerfstyle_b200 import raymarching
    m, a, b = _occ_renderers(dev)
    H, C = a.grid_size, a.cascade
    H3 = H ** 3
    g = torch.Generator(device=dev).manual_seed(11)
    for rnd in range(2):                                        # second round exercises the decay / max against a non-zero grid
        noise = torch.rand(C, H3, 3, device=dev, generator=g)   # Morton order
        # the reference draws its jitter in meshgrid (x, y, z) order: row j = (x * H + y) * H + z  <->  Morton cell i
        ar = torch.arange(H, dtype=torch.int32, device=dev)
        xx, yy, zz = torch.meshgrid(ar, ar, ar, indexing='ij')
        mort = raymarching.morton3D(torch.stack([xx.reshape(-1), yy.reshape(-1), zz.reshape(-1)], dim=-1)).long()
        queue = [noise[c][mort] for c in range(C)]
        monkeypatch.setattr(torch, 'rand_like', lambda t, _q=queue: _q.pop(0).to(t.dtype))
        with torch.autocast('cuda', dtype=torch.float16):
            b.update_state()
        monkeypatch.undo()
        assert not queue
        with torch.autocast('cuda', dtype=torch.float16):
            a.update_state_fused(noise=noise)
        assert torch.equal(a.density_grid, b.density_grid)
        assert float(a.density_grid.max()) > 0
        assert abs(a.mean_density - b.mean_density) <= 1e-6 * abs(b.mean_density)
        diff = int((a.density_bitfield != b.density_bitfield).sum())
        assert diff <= 2, diff
        assert int(a.density_bitfield.count_nonzero()) > 1000
        assert a.mean_count == b.mean_count
        a.local_step = b.local_step = 16                        # still the full phase (< update_thres)


def test_update_state_fused_sparse_phase_semantics(cuda_lib, dev):
    """Later-phase update (renderer.py:157-181) on the device: random cells + random OCCUPIED cells, tmp_grid scatter, decay/max.
    Checked by recomputation from the points the update used: untouched cells keep their value, sampled cells become
    max(old * decay, sigma of a sample in that cell); the occupied picks are occupied; the random cells are morton3D(rnd)."""
    from nerfstyle_b200 import raymarching
    m, a, b = _occ_renderers(dev)
    H, C = a.grid_size, a.cascade
    H3 = H ** 3
    N = H3 // 4
    with torch.autocast('cuda', dtype=torch.float16):
        a.update_state_fused()                                  # populate the grid (full phase)
    a.local_step = a.update_thres + 5
    old = a.density_grid.clone()
    g = torch.Generator(device=dev).manual_seed(3)
    rnd_cells = torch.randint(0, H, (C, N, 3), device=dev, dtype=torch.int32, generator=g)
    pick = torch.rand(C, N, device=dev, generator=g)
    noise = torch.rand(C, 2 * N, 3, device=dev, generator=g)
    with torch.autocast('cuda', dtype=torch.float16):
        a.update_state_fused(noise=noise, rnd_cells=rnd_cells, pick=pick)
        indices, pts = a._occ_last
        sig = (m(pts.view(-1, 3)).reshape(C, 2 * N).float() * a.density_scale)
    for c in range(C):
        assert torch.equal(indices[c, :N], raymarching.morton3D(rnd_cells[c]))
        occ = indices[c, N:].long()
        assert bool((old[c][occ] > 0).all())
        # expected: scatter-amax of the samples, then the reference's masked decay / max
        tmp = torch.full((H3,), -1.0, device=dev).scatter_reduce(0, indices[c].long(), sig[c], reduce='amax', include_self=True)
        exp = torch.where((old[c] >= 0) & (tmp >= 0), torch.maximum(old[c] * a.density_decay, tmp), old[c])
        assert torch.equal(a.density_grid[c], exp)
        assert int((tmp >= 0).sum()) > N // 2 and int((tmp < 0).sum()) > 0
    # the points lie inside their cells' cascade-scaled, jittered boxes
    assert float(pts.abs().max()) <= a.bound
    thr = min(a.mean_density, a.density_thresh)
    ref_bits = raymarching.packbits(a.density_grid, thr)
    assert int((ref_bits != a.density_bitfield).sum()) <= 2


@pytest.mark.parametrize('n_rays,k', [(8192, 8), (1000, 8), (37, 1), (4096, 0)])
def test_fused_recon_loss_matches_torch_composition(cuda_lib, dev, n_rays, k):
    """nrf_recon_loss == white background (renderer.py:229-232) + MSE + class_lambda * CrossEntropy (trainers/base.py:251-304)
    composed from torch ops: loss terms within 1e-6 relative, gradients w.r.t. the compositing outputs within 1e-5 * max."""
    from nerfstyle_b200.trainer import recon_loss
    g = torch.Generator(device=dev).manual_seed(n_rays + k)
    image = (torch.randn(n_rays, 3 + k, device=dev, generator=g) * 2).requires_grad_(True)
    ws = torch.rand(n_rays, device=dev, generator=g).requires_grad_(True)
    tgt = torch.rand(n_rays, 3, device=dev, generator=g)
    cls = torch.randint(0, max(k, 1), (n_rays,), device=dev, generator=g)
    lam = 0.001
    total, mse, ce = recon_loss(ws, image, tgt, cls, lam)
    (total * 1024.0).backward()                                  # an upstream scale, as under GradScaler
    gi, gw = image.grad.clone(), ws.grad.clone()
    image.grad = ws.grad = None
    rgb = image[:, :3] + (1 - ws).unsqueeze(-1)
    e_mse = torch.mean((rgb - tgt) ** 2)
    e_ce = torch.nn.functional.cross_entropy(image[:, 3:], cls) if k > 0 else torch.zeros((), device=dev)
    e_total = e_mse + e_ce * lam
    (e_total * 1024.0).backward()
    assert abs(float(mse) - float(e_mse)) <= 1e-6 * abs(float(e_mse))
    assert abs(float(ce) - float(e_ce)) <= 2e-6 * abs(float(e_ce)) + 1e-12
    assert abs(float(total) - float(e_total)) <= 1e-6 * abs(float(e_total))
    assert float((gi - image.grad).abs().max()) <= 1e-5 * float(image.grad.abs().max())
    assert float((gw - ws.grad).abs().max()) <= 1e-5 * float(ws.grad.abs().max())


def test_trainstep_fused_loss_equals_unfused(cuda_lib, dev):
    """TrainStep(fused_loss=True) and the op-by-op loss run the same trajectory."""
    from nerfstyle_b200 import model as M, scenes
    from nerfstyle_b200.trainer import TrainStep
    intr = dict(scenes.ROOM)
    pose = scenes.synthetic_poses(2, 0)[0]
    losses = []
    for fused in (True, False):
        torch.manual_seed(0)
        torch.cuda.manual_seed_all(0)
        m = M.StyleTCNerf([-2., -2., -2.], [2., 2., 2.], class_dim=8).to(dev)
        r = M.Renderer(m, 2.0, raymarch_channels=11).to(dev)
        ts = TrainStep(r, enable_amp=True, fused_loss=fused)
        gen = torch.Generator().manual_seed(0)
        ls = []
        for it in range(4):
            idx = scenes.frame_indices(intr, 1024, gen).to(dev)
            o, d = scenes.generate_rays(pose, intr, dev, idx)
            tgt, seg = scenes.synthetic_target(idx, intr)
            ls.append(float(ts.step(o, d, tgt, seg)))
        losses.append(ls)
    for a, b in zip(*losses):
        assert abs(a - b) <= 2e-3 * abs(b), losses


def test_occupancy_kernels_against_cpu_oracle(cuda_lib, oracle, dev):
    """nrf_occ_points_full / nrf_occ_update / nrf_packbits_dev against the numpy restatement of renderer.py:120-194
    (oracle/occupancy.py): sample points and updated grid bit-exact, mean within 1e-6, bitfield identical away from the
    threshold.  Small grid (H = 32) so the CPU side takes a fraction of a second."""
    from oracle import occupancy as occ
    from nerfstyle_b200 import _lib as L
    lib = L.lib()
    H, C, bound = 32, 2, 2.0
    H3 = H ** 3
    g = torch.Generator(device=dev).manual_seed(21)
    noise = torch.rand(C, H3, 3, device=dev, generator=g)
    sh = torch.tensor([v for pair in occ.cascade_constants(C, bound, H) for v in pair], dtype=torch.float32, device=dev)
    pts = torch.empty(C, H3, 3, device=dev)
    st = torch.cuda.current_stream().cuda_stream
    assert lib.nrf_occ_points_full(pts.data_ptr(), noise.data_ptr(), H, C, sh.data_ptr(), st) == 0
    e_pts = occ.points_full_morton(noise.cpu().numpy(), C, bound, H)
    assert np.array_equal(pts.cpu().numpy().view(np.uint32), e_pts.view(np.uint32))
    # update: a grid with invalid (-1) cells, sigmas with a spread of magnitudes
    grid = torch.rand(C, H3, device=dev, generator=g) * 4
    grid[torch.rand(C, H3, device=dev, generator=g) < 0.1] = -1.0
    sig = torch.rand(C * H3, device=dev, generator=g) ** 4 * 30
    e_grid, e_mean, e_bits = occ.grid_update(grid.cpu().numpy(), (sig * 1.5).reshape(C, H3).cpu().numpy(), 0.95, 10.0)
    state = torch.zeros(2, device=dev)
    scratch = torch.empty(int(lib.nrf_occ_scratch_bytes()), dtype=torch.uint8, device=dev)
    bits = torch.empty(C * H3 // 8, dtype=torch.uint8, device=dev)
    assert lib.nrf_occ_update(grid.data_ptr(), sig.data_ptr(), 1.5, 0.95, C * H3, 10.0, state.data_ptr(), scratch.data_ptr(), st) == 0
    assert lib.nrf_packbits_dev(grid.data_ptr(), C * H3 // 8, state.data_ptr() + 4, bits.data_ptr(), st) == 0
    assert np.array_equal(grid.cpu().numpy().view(np.uint32), e_grid.view(np.uint32))
    assert abs(float(state[0]) - e_mean) <= 1e-6 * e_mean and abs(float(state[1]) - min(e_mean, 10.0)) <= 1e-6 * e_mean
    assert int((bits.cpu().numpy() != np.asarray(e_bits).reshape(-1)).sum()) <= 1


def test_peer_memory_optimizer_kernels_on_one_gpu(cuda_lib, dev):
    """The kernels of the fused data-parallel exchange (csrc/optim_p2p.cu) with the ranks emulated by several local
    buffers (their "peer pointers" all point into this GPU's memory; no multicast object, so the rank-order load / store
    loops run): nrf_adam_step_pair_p2p over W gradient buffers must equal nrf_adam_step_pair on their sum -- parameters,
    both Adam moments, EMA bit for bit -- and must write the updated fp16 rows into EVERY rank's table copy;
    nrf_small_allreduce_p2p must sum the small buffers in rank order and raise found-inf when any rank flagged it.
    (The multi-process run over NVLink -- multimem.ld_reduce / multimem.st through the NVLS mapping -- is checked by
    tools/check_sharded_optimizer.py under torchrun at N = 2 and N = 8: bit-identical to all-reduce + replicated Adam at N = 2.)"""
    import struct
    from nerfstyle_b200 import _lib as L
    lib = L.lib()
    st = torch.cuda.current_stream().cuda_stream
    g = torch.Generator().manual_seed(5)
    T, W = 40000, 3
    row_lo, rows = 10000, 20000                                   # this "rank" owns rows [10000, 30000) of the pair buffers
    grads = [torch.randn(T, 2, 2, generator=g).to(dev) * 512.0 for _ in range(W)]
    halves = [torch.zeros(T, 2, 2, dtype=torch.float16, device=dev) for _ in range(W)]
    total = grads[0].clone()
    for q in grads[1:]:
        total += q                                                # rank order, like the kernel's loop

    def state(found_inf=0):
        raw = struct.pack('fiii', 512.0, found_inf, 0, 7) + b'\\0' * (int(lib.nrf_opt_state_bytes()) - 16)
        return torch.frombuffer(bytearray(raw), dtype=torch.uint8).to(dev)

    def fresh():
        gg = torch.Generator().manual_seed(6)
        mk = lambda: torch.randn(2 * rows, generator=gg).to(dev)             # noqa: E731  (shard-local tensors)
        return {'p0': mk() * 0.1, 'p1': mk() * 0.1, 'm0': mk() * 0.01, 'm1': mk() * 0.01, 'v0': mk().abs() * 1e-3, 'v1': mk().abs() * 1e-3,
                'e0': mk() * 0.1, 'e1': mk() * 0.1}
    a, b = fresh(), fresh()
    gp = torch.tensor([q.data_ptr() for q in grads], dtype=torch.int64, device=dev)
    hp = torch.tensor([h.data_ptr() for h in halves], dtype=torch.int64, device=dev)
    sa, sb = state(), state()
    args = (0.01, 50.0, 0.9, 0.999, 1e-15, 0.05)
    L.check(lib.nrf_adam_step_pair_p2p(a['p0'].data_ptr(), a['p1'].data_ptr(), gp.data_ptr(), None, hp.data_ptr(), None, W, row_lo,
                                       a['m0'].data_ptr(), a['m1'].data_ptr(), a['v0'].data_ptr(), a['v1'].data_ptr(), a['e0'].data_ptr(),
                                       a['e1'].data_ptr(), rows, sa.data_ptr(), *args, st), 'adam_step_pair_p2p')
    ref_half = torch.zeros(rows, 2, 2, dtype=torch.float16, device=dev)
    L.check(lib.nrf_adam_step_pair(b['p0'].data_ptr(), b['p1'].data_ptr(), total[row_lo:row_lo + rows].contiguous().data_ptr(),
                                   b['m0'].data_ptr(), b['m1'].data_ptr(), b['v0'].data_ptr(), b['v1'].data_ptr(), b['e0'].data_ptr(),
                                   b['e1'].data_ptr(), ref_half.data_ptr(), rows, sb.data_ptr(), *args, st), 'adam_step_pair')
    for k in a:
        assert torch.equal(a[k], b[k]), k
    for h in halves:                                              # the all-gather: every rank's copy got this shard's rows, nothing else
        assert torch.equal(h[row_lo:row_lo + rows], ref_half)
        assert float(h[:row_lo].abs().max()) == 0.0 and float(h[row_lo + rows:].abs().max()) == 0.0
    assert float(ref_half.abs().max()) > 0
    # found-inf on the state block skips the update (and the table writes) like nrf_adam_step_pair
    c, sc = fresh(), state(found_inf=1)
    before = {k: v.clone() for k, v in c.items()}
    halves2 = [torch.zeros(T, 2, 2, dtype=torch.float16, device=dev)]
    hp2 = torch.tensor([halves2[0].data_ptr()], dtype=torch.int64, device=dev)
    L.check(lib.nrf_adam_step_pair_p2p(c['p0'].data_ptr(), c['p1'].data_ptr(), gp.data_ptr(), None, hp2.data_ptr(), None, 1, row_lo,
                                       c['m0'].data_ptr(), c['m1'].data_ptr(), c['v0'].data_ptr(), c['v1'].data_ptr(), c['e0'].data_ptr(),
                                       c['e1'].data_ptr(), rows, sc.data_ptr(), *args, st), 'adam_step_pair_p2p')
    for k in ('p0', 'p1', 'm0', 'm1', 'v0', 'v1'):
        assert torch.equal(c[k], before[k]), k
    assert float(halves2[0].abs().max()) == 0.0
    # ---- small tensors + flag
    n = 3077
    bufs = [torch.randn(n + 8, generator=g).to(dev) for _ in range(W)]
    for q in bufs:
        q[n:] = 0.0
    sp = torch.tensor([q.data_ptr() for q in bufs], dtype=torch.int64, device=dev)
    out = torch.empty(n, device=dev)
    s0 = state()
    L.check(lib.nrf_small_allreduce_p2p(sp.data_ptr(), W, n, out.data_ptr(), s0.data_ptr(), st), 'small_allreduce_p2p')
    want = bufs[0][:n].clone()
    for q in bufs[1:]:
        want += q[:n]
    assert torch.equal(out, want)
    assert int(s0[4:8].view(torch.int32)) == 0
    bufs[1][n] = 1.0                                              # rank 1 saw an inf
    L.check(lib.nrf_small_allreduce_p2p(sp.data_ptr(), W, n, out.data_ptr(), s0.data_ptr(), st), 'small_allreduce_p2p')
    assert int(s0[4:8].view(torch.int32)) == 1
