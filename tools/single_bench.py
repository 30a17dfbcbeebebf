"""Single-table hash-grid scatter (the reference-facing GridEncoder's backward): walk form vs thread-per-sample kernel.
Usage (GPU box): python tools/single_bench.py"""
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
for half in (True, False):
    dt = 1 if half else 0
    grad = torch.randn(B, 32, device=dev).to(torch.float16 if half else torch.float32)
    ge = torch.zeros(enc.embeddings.shape, dtype=torch.float32, device=dev)
    f = lambda: lib.nrf_grid_encode_backward(grad.data_ptr(), pts.data_ptr(), None, enc.offsets.data_ptr(), ge.data_ptr(), B, 3, 2, 16, S,
                                             16, 0, None, None, 0, 1, 0, dt, 0, 1, st)
    res = {}
    for walk in (0, 64, 128, 256):
        lib.nrf_grid_set_bwd_walk(walk)
        ge.zero_()
        assert f() == 0
        torch.cuda.synchronize()
        res[walk] = ge.clone()
        print('single-table bwd half=%d walk=%3d: %.3f ms (%d points)' % (half, walk, mb.timeit(f), B))
    lib.nrf_grid_set_bwd_walk(128)
    print('   walk vs thread-per-sample: rel max err %.1e' % float((res[128] - res[0]).abs().max() / res[0].abs().max()))
