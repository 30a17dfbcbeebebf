"""Per-kernel micro-benchmarks on a realistic sample set (8192 room-shaped rays through the analytic occupancy).
Usage (GPU box): python tools/microbench.py [grid|mlp|march|composite|all]"""
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from nerfstyle_b200 import _lib, model as M, raymarching, scenes, tcnn  # noqa: E402


dev = torch.device('cuda:0')
lib = _lib.lib()


def timeit(fn, iters=10, warm=3):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(iters):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    ts.sort()
    return ts[len(ts) // 2]


def make_samples(n_rays=8192, kind='bernoulli'):
    intr = dict(scenes.ROOM)
    pose = scenes.synthetic_poses(4, 0)[1]
    gen = torch.Generator().manual_seed(0)
    idx = scenes.frame_indices(intr, n_rays, gen).to(dev)
    o, d = scenes.generate_rays(pose, intr, dev, idx)
    grid = scenes.bernoulli_density_grid(2, 128, 0.5, 1) if kind == 'bernoulli' else scenes.analytic_density_grid(2, 128, 2.0)
    bits = raymarching.packbits(grid.to(dev), 0.5)
    aabb = torch.tensor([-2., -2, -2, 2, 2, 2], device=dev)
    nears, fars = raymarching.near_far_from_aabb(o, d, aabb, 0.2)
    return o, d, bits, nears, fars


def bench_march():
    o, d, bits, nears, fars = make_samples()
    def f():
        c = torch.zeros(2, dtype=torch.int32, device=dev)
        return raymarching.march_rays_train(o, d, None, 2.0, bits, 2, 128, nears, fars, c, -1, False, 128, True, 0., 1024, False)
    for mode in (0, 1):
        lib.nrf_march_set_mode(mode)
        print('march mode %d (0 thread-per-ray, 1 warp-per-ray): %.3f ms' % (mode, timeit(f)))
    xyzs, dirs, deltas, rays = f()
    print('march_rays_train (count+sync+write): %.3f ms for %d rays, %d samples' % (timeit(f), o.shape[0], xyzs.shape[0]))
    N = o.shape[0]
    scratch = _lib.scratch(dev, lib.nrf_march_scratch_bytes(N))
    c = torch.zeros(2, dtype=torch.int32, device=dev)
    st = torch.cuda.current_stream().cuda_stream
    r2 = torch.empty_like(rays)
    t = timeit(lambda: lib.nrf_march_rays_train_count(o.data_ptr(), d.data_ptr(), bits.data_ptr(), 2.0, 0.0, 1024, N, 2, 128,
                                                      nears.data_ptr(), fars.data_ptr(), None, r2.data_ptr(), c.data_ptr(),
                                                      scratch.data_ptr(), st))
    print('  count pass: %.3f ms' % t)
    M_ = xyzs.shape[0]
    t = timeit(lambda: lib.nrf_march_rays_train_write(o.data_ptr(), d.data_ptr(), None, bits.data_ptr(), 2.0, 0.0, 1024, 0, N, 2, 128,
                                                      N * 1024, M_, M_, nears.data_ptr(), fars.data_ptr(), None, rays.data_ptr(),
                                                      xyzs.data_ptr(), dirs.data_ptr(), deltas.data_ptr(), st))
    print('  write pass: %.3f ms  (%.1f GB/s of 40 B/sample)' % (t, M_ * 40 / t / 1e6))
    return xyzs, dirs, deltas, rays


def bench_grid(xyzs):
    enc = M.get_grid_encoder(max_bound=4.0).to(dev)
    pts = ((xyzs + 2.0) / 4.0 + 1) / 2          # model normalisation + the [0.5,1] octant quirk
    B = pts.shape[0]
    S = float(np.float32(np.log2(enc.per_level_scale)))
    st = torch.cuda.current_stream().cuda_stream
    for half in (True, False):
        emb = enc.embeddings.detach().half() if half else enc.embeddings.detach()
        dt = 1 if half else 0
        out = torch.empty(B, 32, dtype=emb.dtype, device=dev)
        grad = torch.randn(B, 32, device=dev).to(emb.dtype)
        ge = torch.zeros(enc.embeddings.shape, dtype=torch.float32, device=dev)
        fb = 588 if half else 1164
        bb = 1100 if half else 2188
        for lpt in (16, 8, 4, 2, 1):
            lib.nrf_grid_set_tuning(lpt, 0, -1)
            t = timeit(lambda: lib.nrf_grid_encode_forward(pts.data_ptr(), emb.data_ptr(), enc.offsets.data_ptr(), out.data_ptr(), B, 3,
                                                           2, 16, S, 16, 0, None, 0, 1, 0, dt, 1, st))
            print('grid fwd  half=%d lpt=%2d: %.3f ms  %.0f GB/s algorithmic' % (half, lpt, t, B * fb / t / 1e6))
        lib.nrf_grid_set_tuning(16, 0, -1)
        for lpt in (16, 4, 1):
            for agg in (0, 1):
                lib.nrf_grid_set_tuning(0, lpt, agg)
                t = timeit(lambda: lib.nrf_grid_encode_backward(grad.data_ptr(), pts.data_ptr(), None, enc.offsets.data_ptr(),
                                                                ge.data_ptr(), B, 3, 2, 16, S, 16, 0, None, None, 0, 1, 0, dt, 0, 1, st))
                print('grid bwd  half=%d lpt=%2d agg=%2d: %.3f ms  %.0f GB/s algorithmic' % (half, lpt, agg, t, B * bb / t / 1e6))
        lib.nrf_grid_set_tuning(0, 16, 1)
        # dual (two tables, one index computation) and the paired 16-byte reductions
        emb2, out2, ge2, grad2 = emb.clone(), torch.empty_like(out), torch.zeros_like(ge), grad.clone()
        t = timeit(lambda: lib.nrf_grid_encode_forward_dual(pts.data_ptr(), emb.data_ptr(), emb2.data_ptr(), enc.offsets.data_ptr(),
                                                            out.data_ptr(), out2.data_ptr(), B, 16, S, 16, 0, 1, 0, dt, None, st))
        print('grid fwd DUAL half=%d: %.3f ms for two encoders  (%.0f GB/s algorithmic)' % (half, t, 2 * B * fb / t / 1e6))
        for pr in (0,):
            t1 = timeit(lambda: lib.nrf_grid_encode_backward(grad.data_ptr(), pts.data_ptr(), None, enc.offsets.data_ptr(),
                                                             ge.data_ptr(), B, 3, 2, 16, S, 16, 0, None, None, 0, 1, 0, dt, 0, 1, st))
            t2 = timeit(lambda: lib.nrf_grid_encode_backward_dual(grad.data_ptr(), grad2.data_ptr(), pts.data_ptr(), enc.offsets.data_ptr(),
                                                                  ge.data_ptr(), ge2.data_ptr(), B, 16, S, 16, 0, 1, 0, dt, 0, None, st))
            print('grid bwd  half=%d (%d): single %.3f ms (%.0f GB/s)   DUAL %.3f ms for two encoders (%.0f GB/s)' % (
                half, pr, t1, B * bb / t1 / 1e6, t2, 2 * B * bb / t2 / 1e6))


def bench_mlp(B):
    for name, (ni, no, nh, act) in {'density': (32, 1, 1, 'None'), 'class': (32, 8, 1, 'None'), 'color1': (32, 16, 1, 'None'),
                                    'color2': (16, 3, 2, 'Sigmoid')}.items():
        net = tcnn.Network(ni, no, {'otype': 'FullyFusedMLP', 'activation': 'ReLU', 'output_activation': act, 'n_neurons': 64,
                                    'n_hidden_layers': nh}).to(dev)
        x = torch.randn(B, ni, device=dev).half().requires_grad_(True)
        dy = torch.randn(B, no, device=dev).half()
        with torch.no_grad():
            tf = timeit(lambda: net(x))
        y = net(x)
        tb = timeit(lambda: torch.autograd.grad(y, [x, net.params], dy, retain_graph=True))
        fl = 2 * sum(r * c for r, c in net.layer_shapes) * B
        print('mlp %-8s B=%d: fwd %.3f ms (%.1f TFLOP/s, %.0f GB/s io)  bwd %.3f ms (%.1f TFLOP/s)' % (
            name, B, tf, fl / tf / 1e9, B * (ni + no) * 2 / tf / 1e6, tb, 2 * fl / tb / 1e9))


def bench_composite(deltas, rays):
    Mrows = deltas.shape[0]
    sig = (torch.rand(Mrows, device=dev) * 2).requires_grad_(True)
    rgb = torch.rand(Mrows, 11, device=dev).requires_grad_(True)
    with torch.no_grad():
        tf = timeit(lambda: raymarching.composite_rays_train(sig, rgb, deltas, rays, 1e-4, False))
    ws, depth, image = raymarching.composite_rays_train(sig, rgb, deltas, rays, 1e-4, False)
    gw, gi = torch.randn_like(ws), torch.randn_like(image)
    tb = timeit(lambda: torch.autograd.grad([ws, image], [sig, rgb], [gw, gi], retain_graph=True))
    print('composite fwd %.3f ms (%.0f GB/s of 64 B/sample)  bwd %.3f ms incl. zero fills (%.0f GB/s of 112 B/sample)' % (
        tf, Mrows * 64 / tf / 1e6, tb, Mrows * 112 / tb / 1e6))


if __name__ == '__main__':
    what = sys.argv[1] if len(sys.argv) > 1 else 'all'
    xyzs, dirs, deltas, rays = bench_march()
    if what in ('grid', 'all'):
        bench_grid(xyzs)
    if what in ('mlp', 'all'):
        bench_mlp(xyzs.shape[0])
    if what in ('composite', 'all'):
        bench_composite(deltas, rays)


def bench_update_state():
    import bench as B
    for fused in (True, False):
        from nerfstyle_b200.trainer import TrainStep
        ts = B.build_trainer(dev, True, 1)
        if not fused:
            ts = TrainStep(ts.renderer, enable_amp=True, world_size=1, fused_optimizer=False)
        host, devb = B.make_batches(40, 8192, 0, 1, dev)
        r = ts.renderer
        with torch.autocast('cuda', dtype=torch.float16):
            t_up = timeit(lambda: r.update_state(), iters=5, warm=2)
        r.update_occ = True
        for s in range(4):
            ts.step(*B.unpack(devb[s]))
        torch.cuda.synchronize()
        # steps without occupancy updates
        r.update_occ = False
        t0 = time.perf_counter()
        for s in range(4, 36):
            ts.step(*B.unpack(devb[s]))
        torch.cuda.synchronize()
        t_step = (time.perf_counter() - t0) / 32 * 1e3
        print('fused_optimizer=%s: update_state %.2f ms per call; train step without update %.3f ms' % (fused, t_up, t_step))


if __name__ == '__main__' and len(sys.argv) > 1 and sys.argv[1] == 'update':
    bench_update_state()
