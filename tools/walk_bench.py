"""The paired hash-grid scatter alone on the bench's sample set (walk form), for ncu.  Usage: python tools/walk_bench.py [chunk] [queue]"""
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
chunk = int(sys.argv[1]) if len(sys.argv) > 1 else 128
queue = int(sys.argv[2]) if len(sys.argv) > 2 else 1
xyzs, dirs, deltas, rays = mb.bench_march()
enc = M.get_grid_encoder(max_bound=4.0).to(dev)
pts = ((xyzs + 2.0) / 4.0 + 1) / 2
B = pts.shape[0]
S = float(np.float32(np.log2(enc.per_level_scale)))
st = torch.cuda.current_stream().cuda_stream
g0 = torch.randn(B, 32, device=dev).half()
g1 = torch.randn(B, 32, device=dev).half()
gp = torch.zeros(enc.embeddings.shape[0], 2, 2, dtype=torch.float32, device=dev)
lib.nrf_grid_set_bwd_walk(chunk)
lib.nrf_grid_set_bwd_walk_queue(queue)
f = lambda: lib.nrf_grid_encode_backward_pair(g0.data_ptr(), g1.data_ptr(), pts.data_ptr(), enc.offsets.data_ptr(), gp.data_ptr(),
                                              B, 16, S, 16, 0, 1, 0, 1, None, st)
assert f() == 0
torch.cuda.synchronize()
print('walk %d queue %d: %.3f ms (%d points)' % (chunk, queue, mb.timeit(f), B))
