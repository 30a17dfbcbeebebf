"""Checkpoint round trips (SURVEY.md 8f NEXT-4; renderer.py:78-107, trainers/base.py:231-249) and the fused optimizer's
fp16 shadow copies across them: the reference builds its optimizer BEFORE load_ckpt and evaluates under EMA
weights, so a state loaded after the optimizer exists -- or swapped in by the EMA scope -- must be what the next render
reads, in every dtype the kernels consume."""
import numpy as np
import pytest
import torch

import refenv

pytestmark = pytest.mark.gpu
BOUND, K = 2.0, 8


def _stack(dev, seed):
    from nerfstyle_b200 import model as M
    torch.manual_seed(seed)
    m = M.StyleTCNerf([-BOUND] * 3, [BOUND] * 3, class_dim=K).to(dev)
    with torch.no_grad():
        for e in (m.x_density_embedder, m.x_color_embedder):
            e.embeddings.uniform_(-0.5, 0.5)
    r = M.Renderer(m, BOUND, raymarch_channels=3 + K).to(dev)
    return m, r


def _train(r, dev, steps, seed=0, **kw):
    from nerfstyle_b200 import scenes
    from nerfstyle_b200.trainer import TrainStep
    ts = TrainStep(r, **kw)
    g = torch.Generator().manual_seed(seed)
    for i in range(steps):
        o, d = scenes.random_rays(1024, 100 + i, dev)
        ts.step(o, d, torch.rand(1024, 3, generator=g).to(dev), torch.randint(0, K, (1024,), generator=g).to(dev))
    return ts


def _frame(r, dev, graph=False):
    from nerfstyle_b200 import scenes
    o, d = scenes.random_rays(4096, 7, dev)
    with torch.no_grad(), torch.autocast('cuda', dtype=torch.float16):
        return r.render_test_graph(o, d) if graph else r.render_test(o, d)


def test_state_dict_round_trip_with_optimizer_built_first(cuda_lib, dev):
    m, r = _stack(dev, 1)
    _train(r, dev, 3)
    sd = r.state_dict()
    assert sorted(sd.keys()) == ['bound', 'density_bitfield', 'density_grid', 'local_step', 'mean_count', 'mean_density', 'model',
                                 'raymarch_channels', 'step_counter']
    # exactly the reference model's keys (no bbox buffers): a strict load into / from the reference's StyleTCNerf works
    assert sorted(sd['model'].keys()) == ['class_net.params', 'color1_net.params', 'color2_net.params', 'density_net.params',
                                          'x_color_embedder.embeddings', 'x_color_embedder.offsets',
                                          'x_density_embedder.embeddings', 'x_density_embedder.offsets']
    sd = {k: (v.clone() if torch.is_tensor(v) else v) for k, v in sd.items()}
    sd['model'] = {k: v.clone() for k, v in sd['model'].items()}
    want = _frame(r, dev)
    # a fresh stack whose optimizer (fp16 shadows, interleaved pair buffer) exists BEFORE the checkpoint is loaded
    m2, r2 = _stack(dev, 2)
    ts2 = _train(r2, dev, 1, seed=5)
    before = _frame(r2, dev)
    assert not torch.equal(before[0], want[0])
    r2.load_state_dict(sd)
    assert r2.local_step == r.local_step and r2.mean_count == r.mean_count
    got = _frame(r2, dev)
    for a, b in zip(want, got):
        assert torch.equal(a, b)                      # same weights, same bitfield, same kernels: bit-identical
    got_graph = _frame(r2, dev, graph=True)           # the CUDA-graph loop re-reads tables / weights per frame
    assert float((got_graph[0] - want[0]).abs().max()) < 2e-3
    # training continues from the loaded weights (the fp16 copies the kernels read are the loaded ones)
    for p, q in zip(m.parameters(), m2.parameters()):
        assert torch.equal(p, q)
    assert torch.equal(ts2.fused.half_pair[:, 0].float(), m.x_density_embedder.embeddings.detach().half().float())


def test_reference_renderer_loads_our_checkpoint_and_back(cuda_lib, dev):
    """The reference's own Renderer.load_state_dict / state_dict (renderer.py:78-107) against this package's: a checkpoint
    written by either side renders the same frame on the other."""
    from nerfstyle_b200 import scenes
    m, r = _stack(dev, 3)
    _train(r, dev, 2)
    sd = r.state_dict()
    o, d = scenes.random_rays(2048, 9, dev)
    with torch.no_grad(), torch.autocast('cuda', dtype=torch.float16):
        want = r.render_test(o, d, sync_every=1)
    with refenv.ReferenceEnv() as E:
        E.nerf_lib.nerf_lib.device = dev
        bbox = E.common.BBox.from_radius(BOUND)
        model = E.style_nerf.StyleTCNerf(E.network_config(), bbox, K, torch.float16, use_dir=False)
        intr = E.common.Intrinsics(378, 504, 383.83, 383.83, 252., 189.)
        rr = E.renderer.Renderer(model, E.renderer_config(), intr, BOUND, raymarch_channels=3 + K).to(dev)
        rr.load_state_dict(dict(sd, intr=intr, precrop_frac=1.))
        with torch.no_grad(), torch.autocast('cuda', dtype=torch.float16):
            got = rr.render_test(E.common.RayBatch(o, d))
        for a, b in zip(want, got):
            assert float((a - b).abs().max()) < 2e-3      # fused heads (ours) vs the reference's op sequence
        back = rr.state_dict()
    m3, r3 = _stack(dev, 4)
    r3.load_state_dict(back)
    with torch.no_grad(), torch.autocast('cuda', dtype=torch.float16):
        again = r3.render_test(o, d, sync_every=1)
    for a, b in zip(want, again):
        assert torch.equal(a, b)


def test_ema_scope_renders_with_ema_weights(cuda_lib, dev):
    m, r = _stack(dev, 5)
    ts = _train(r, dev, 4)
    train_frame = _frame(r, dev)
    ema = [e.clone() for e in ts.fused.ema]
    with ts.fused.ema_scope():
        ema_frame = _frame(r, dev)
        for p, e in zip(m.parameters(), ema):
            assert torch.equal(p.detach(), e)
    assert not torch.equal(ema_frame[0], train_frame[0])
    assert torch.equal(_frame(r, dev)[0], train_frame[0])          # training weights (and their fp16 copies) are back
    # the same frame from a plain model holding the EMA weights
    m2, r2 = _stack(dev, 6)
    with torch.no_grad():
        for p, e in zip(m2.parameters(), ema):
            p.copy_(e)
    r2.density_bitfield.copy_(r.density_bitfield)
    assert torch.equal(_frame(r2, dev)[0], ema_frame[0])


def test_second_optimizer_takes_over_cleanly(cuda_lib, dev):
    """trainers/base.py:_reset_optim builds a new optimizer between training stages: the tables must keep training."""
    m, r = _stack(dev, 7)
    ts1 = _train(r, dev, 2)
    before = m.x_color_embedder.embeddings.detach().clone()
    ts2 = _train(r, dev, 2, seed=3, fused_optimizer=False)         # torch.optim.Adam on ordinary .grad
    assert not ts1.fused.alive
    assert not hasattr(m.x_color_embedder.embeddings, '_nrf_grad_sink')
    assert m.x_color_embedder.embeddings.grad is not None
    assert float((m.x_color_embedder.embeddings.detach() - before).abs().max()) > 0
    ts3 = _train(r, dev, 1, seed=4)                                # and back to the fused one
    assert ts3.fused.alive and ts3.fused.grad_pair is not None
