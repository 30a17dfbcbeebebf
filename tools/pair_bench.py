"""Dual vs paired (interleaved) hash-grid kernels on the bench's sample set: values must be identical, times are printed.
Usage (GPU box): python tools/pair_bench.py"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, 'tools'))
import microbench as mb  # noqa: E402
from nerfstyle_b200 import model as M  # noqa: E402

dev, lib = mb.dev, mb.lib
xyzs, dirs, deltas, rays = mb.bench_march()
enc = M.get_grid_encoder(max_bound=4.0).to(dev)
pts = ((xyzs + 2.0) / 4.0 + 1) / 2
B = pts.shape[0]
S = float(np.float32(np.log2(enc.per_level_scale)))
st = torch.cuda.current_stream().cuda_stream
off = enc.offsets
for half in (True, False):
    dtc = 1 if half else 0
    dt = torch.float16 if half else torch.float32
    t0 = (torch.rand(enc.embeddings.shape, device=dev) * 2 - 1).to(dt)
    t1 = (torch.rand(enc.embeddings.shape, device=dev) * 2 - 1).to(dt)
    pair = torch.stack([t0, t1], dim=1).contiguous()            # [rows, 2, 2]
    o0, o1, p0, p1 = (torch.empty(B, 32, dtype=dt, device=dev) for _ in range(4))
    f_dual = lambda: lib.nrf_grid_encode_forward_dual(pts.data_ptr(), t0.data_ptr(), t1.data_ptr(), off.data_ptr(), o0.data_ptr(),
                                                      o1.data_ptr(), B, 16, S, 16, 0, 1, 0, dtc, None, st)
    f_pair = lambda: lib.nrf_grid_encode_forward_pair(pts.data_ptr(), pair.data_ptr(), off.data_ptr(), p0.data_ptr(), p1.data_ptr(),
                                                      B, 16, S, 16, 0, 1, 0, dtc, None, None, None, st)
    assert f_dual() == 0 and f_pair() == 0
    torch.cuda.synchronize()
    assert torch.equal(o0, p0) and torch.equal(o1, p1), 'paired forward differs'
    td, tp = mb.timeit(f_dual), mb.timeit(f_pair)
    print('fwd half=%d  dual %.3f ms   pair %.3f ms  (%d points)' % (half, td, tp, B))
    g0 = torch.randn(B, 32, device=dev).to(dt)
    g1 = torch.randn(B, 32, device=dev).to(dt)
    ge0 = torch.zeros(enc.embeddings.shape, dtype=torch.float32, device=dev)
    ge1 = torch.zeros_like(ge0)
    gp = torch.zeros(enc.embeddings.shape[0], 2, 2, dtype=torch.float32, device=dev)
    b_dual = lambda: lib.nrf_grid_encode_backward_dual(g0.data_ptr(), g1.data_ptr(), pts.data_ptr(), off.data_ptr(), ge0.data_ptr(),
                                                       ge1.data_ptr(), B, 16, S, 16, 0, 1, 0, dtc, 0, None, st)
    b_pair = lambda: lib.nrf_grid_encode_backward_pair(g0.data_ptr(), g1.data_ptr(), pts.data_ptr(), off.data_ptr(), gp.data_ptr(),
                                                       B, 16, S, 16, 0, 1, 0, dtc, None, st)
    lib.nrf_grid_set_bwd_walk(0)                                # the thread-per-sample form first
    assert b_dual() == 0 and b_pair() == 0
    torch.cuda.synchronize()
    e0 = float((gp[:, 0] - ge0).abs().max() / ge0.abs().max())
    e1 = float((gp[:, 1] - ge1).abs().max() / ge1.abs().max())
    print('bwd half=%d  paired vs dual gradient: rel max err %.2e / %.2e (atomic order)' % (half, e0, e1))
    assert e0 < 1e-5 and e1 < 1e-5
    td, tp = mb.timeit(b_dual), mb.timeit(b_pair)
    print('bwd half=%d  dual %.3f ms   pair %.3f ms' % (half, td, tp))
    # long runs through the shared-memory transpose instead of the shuffle butterfly: threshold sweep + value check
    ge0.zero_(); ge1.zero_()
    assert b_dual() == 0
    for tr in (1 << 30, 17, 9, 5, 3, 2):
        lib.nrf_grid_set_transpose_min(tr)
        gp.zero_()
        assert b_pair() == 0
        torch.cuda.synchronize()
        e0 = float((gp[:, 0] - ge0).abs().max() / ge0.abs().max())
        e1 = float((gp[:, 1] - ge1).abs().max() / ge1.abs().max())
        assert e0 < 1e-5 and e1 < 1e-5, (tr, e0, e1)
        print('bwd half=%d  pair, transpose for runs >= %-10d: %.3f ms   (rel err vs dual %.1e / %.1e)' % (half, tr, mb.timeit(b_pair), e0, e1))
    lib.nrf_grid_set_transpose_min(9)
    for amin in (1, 2, 3, 4, 5):
        lib.nrf_grid_set_tuning(0, 0, amin)
        print('bwd half=%d  pair, aggregate only when the longest run >= %d: %.3f ms' % (half, amin, mb.timeit(b_pair)))
    lib.nrf_grid_set_tuning(0, 0, 1)
    # the walk form (one lane per chunk of consecutive samples and level, current cell in registers)
    for queue in (0, 1):
        lib.nrf_grid_set_bwd_walk_queue(queue)
        for ch in ((16, 32, 64) if not queue else (32, 64, 128, 256)):
            lib.nrf_grid_set_bwd_walk(ch)
            gp.zero_()
            assert b_pair() == 0
            torch.cuda.synchronize()
            e0 = float((gp[:, 0] - ge0).abs().max() / ge0.abs().max())
            e1 = float((gp[:, 1] - ge1).abs().max() / ge1.abs().max())
            assert e0 < 1e-5 and e1 < 1e-5, (ch, e0, e1)
            print('bwd half=%d  pair, walk chunks of %d, queue %d: %.3f ms   (rel err vs dual %.1e / %.1e)' % (half, ch, queue, mb.timeit(b_pair), e0, e1))
    lib.nrf_grid_set_bwd_walk(128)
    lib.nrf_grid_set_bwd_walk_queue(1)
    ge0.zero_(); ge1.zero_(); gp.zero_()
    assert b_dual() == 0 and b_pair() == 0
    torch.cuda.synchronize()
    e0 = float((gp[:, 0] - ge0).abs().max() / ge0.abs().max())
    print('bwd half=%d  dual as two single-table walks: %.3f ms   (rel err vs pair %.1e)' % (half, mb.timeit(b_dual), e0))
