"""Measurements for BASELINE.json configs 3-5 (bench.py covers config 2, tests cover config 1).

    python tools/bench_configs.py render  [--w 1008 --h 756 --frames 3]        # config 3 (torchrun for N>1: tiles shard)
    python tools/bench_configs.py nnfm                                         # config 4: matching kernel vs torch composition
    python tools/bench_configs.py sweep                                        # config 5: rays/step sweep (torchrun for N>1)
Each prints one JSON line.
"""
import argparse
import json
import os
import sys
import time

import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from nerfstyle_b200 import model as M, nnfm, parallel, raymarching, scenes  # noqa: E402
import bench as B  # noqa: E402


def setup():
    world = int(os.environ.get('WORLD_SIZE', '1'))
    rank = int(os.environ.get('RANK', '0'))
    local = int(os.environ.get('LOCAL_RANK', '0'))
    torch.cuda.set_device(local)
    dev = torch.device('cuda', local)
    if world > 1:
        dist.init_process_group('nccl', device_id=dev)
    return world, rank, dev


def sync(world):
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()


def render(args):
    world, rank, dev = setup()
    torch.manual_seed(0)
    m = M.StyleTCNerf([-2., -2., -2.], [2., 2., 2.], class_dim=B.N_CLASSES).to(dev)
    r = M.Renderer(m, 2.0, raymarch_channels=3 + B.N_CLASSES, density_scale=args.density_scale).to(dev)
    if args.occupancy == 'field':
        with torch.autocast('cuda', dtype=torch.float16):
            r.update_state()
    else:
        r.density_bitfield = raymarching.packbits(scenes.analytic_density_grid(2, 128, 2.0).to(dev), 0.5)
    intr = scenes.scaled_intrinsics(args.w, args.h)
    poses = scenes.synthetic_poses(8, 1)
    n_total = args.w * args.h
    lo, hi = parallel.shard_bounds(n_total, rank, world)
    idx = torch.arange(lo, hi, device=dev)
    times = []
    for f in range(args.frames + 1):
        o, d = scenes.generate_rays(poses[f % len(poses)], intr, dev, idx)
        sync(world)
        t0 = time.perf_counter()
        with torch.no_grad(), torch.autocast('cuda', dtype=torch.float16):
            if args.loop == 'graph':
                img, depth, cls = r.render_test_graph(o, d)
            else:
                img, depth, cls = r.render_test(o, d, sync_every=args.sync_every)
        full = parallel.gather_rows(torch.cat([img, depth[:, None], cls], dim=1), n_total, rank, world)
        sync(world)
        if f > 0:
            times.append(time.perf_counter() - t0)
    t = torch.tensor([sum(times) / len(times)], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    if rank == 0:
        sec = float(t.item())
        print(json.dumps({'config': 'render_full_frame', 'w': args.w, 'h': args.h, 'n_gpus': world, 'rays_per_frame': n_total,
                          'ms_per_frame': round(sec * 1e3, 2), 'mrays_per_s': round(n_total / sec / 1e6, 3),
                          'occupancy': args.occupancy, 'density_scale': args.density_scale, 'frames': args.frames,
                          'sync_every': args.sync_every, 'loop': args.loop,
                          'gathered_bytes': int(n_total * (4 + B.N_CLASSES) * 4)}))
    if world > 1:
        dist.destroy_process_group()


def nnfm_bench(args):
    world, rank, dev = setup()
    from oracle import matching as om      # the reference composition (torch ops) timed beside the kernel
    N1, N2, K = 11844, 15876, 768
    g = torch.Generator().manual_seed(0)
    a = torch.randn(N1, K, generator=g).to(dev)
    b = torch.randn(N2, K, generator=g).to(dev)
    a_hat = (a / a.norm(dim=1, keepdim=True)).half()
    b_hat = (b / b.norm(dim=1, keepdim=True)).half()
    preds = torch.randint(0, 8, (N1,), generator=g).to(dev)
    clusters = (torch.arange(N2) * 8 // N2).to(dev)
    match = list(range(8))

    from nerfstyle_b200 import _lib

    def ours():
        _lib.lib().nrf_nnfm_set_mode(0)
        return nnfm.nn_match(a_hat, b_hat, preds, clusters, match)

    def ours_mma_sync():
        _lib.lib().nrf_nnfm_set_mode(1)
        out = nnfm.nn_match(a_hat, b_hat, preds, clusters, match)
        _lib.lib().nrf_nnfm_set_mode(0)
        return out

    def ref():
        with torch.autocast('cuda', dtype=torch.float16):
            return om.semantic_nn_loss(a, b, preds, clusters, match, 8)
    res = {}
    for name, fn in (('ours', ours), ('ours_mma_sync', ours_mma_sync), ('torch_composition', ref)):
        for _ in range(3):
            fn()
        torch.cuda.synchronize()
        torch.cuda.reset_peak_memory_stats()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(10):
            fn()
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / 10
        res[name] = {'ms': round(ms, 3), 'tflops': round(2.0 * N1 * N2 * K / ms / 1e9, 1),
                     'peak_mem_mb': round(torch.cuda.max_memory_allocated() / 2 ** 20, 1)}
    print(json.dumps({'config': 'nnfm_matching', 'N1': N1, 'N2': N2, 'K': K, **res}))


def sweep(args):
    world, rank, dev = setup()
    out = []
    for log2n in range(14, 21):
        n_global = 1 << log2n
        n_local = n_global // world
        # ~520 samples per room-shaped ray at random init, ~600 B of activations + gradients per sample
        need = n_local * 520 * 600
        free, _ = torch.cuda.mem_get_info(dev)
        if need > 0.6 * free:
            out.append({'rays_per_step': n_global, 'skipped': 'needs ~%d GB per GPU' % (need >> 30)})
            continue
        ts = B.build_trainer(dev, True, world)
        host, devb = B.make_batches(10, n_local, rank, world, dev)
        assert devb[0].shape[0] == n_local
        for s in range(4):
            ts.step(*B.unpack(devb[s]), n_global=n_global)
        sync(world)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for s in range(4, 10):
            ts.step(*B.unpack(devb[s]), n_global=n_global)
        e1.record()
        sync(world)
        t = torch.tensor([e0.elapsed_time(e1) / 6], device=dev, dtype=torch.float64)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        out.append({'rays_per_step': n_global, 'ms_per_step': round(float(t.item()), 3),
                    'rays_per_s': round(n_global / float(t.item()) * 1e3, 1)})
        del ts, host, devb
        torch.cuda.empty_cache()
    if rank == 0:
        print(json.dumps({'config': 'ray_batch_sweep', 'n_gpus': world, 'points': out}))
    if world > 1:
        dist.destroy_process_group()


if __name__ == '__main__':
    ap = argparse.ArgumentParser()
    ap.add_argument('what', choices=['render', 'nnfm', 'sweep'])
    ap.add_argument('--w', type=int, default=1008)
    ap.add_argument('--h', type=int, default=756)
    ap.add_argument('--frames', type=int, default=3)
    ap.add_argument('--occupancy', default='field', choices=['field', 'analytic'])
    ap.add_argument('--density-scale', type=float, default=1.0)
    ap.add_argument('--sync-every', type=int, default=4)
    ap.add_argument('--loop', default='graph', choices=['graph', 'host'])
    a = ap.parse_args()
    {'render': render, 'nnfm': nnfm_bench, 'sweep': sweep}[a.what](a)
