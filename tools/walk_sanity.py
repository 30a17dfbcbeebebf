"""Small run of every walk-form scatter variant (paired ring / paired register-flush / single-table ring), for compute-sanitizer.
Usage (GPU box): python tools/walk_sanity.py   |   compute-sanitizer --tool memcheck|racecheck python tools/walk_sanity.py"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from nerfstyle_b200 import _lib  # noqa: E402
from nerfstyle_b200 import model as M  # noqa: E402

dev = torch.device('cuda:0')
lib = _lib.lib()
enc = M.get_grid_encoder(max_bound=4.0).to(dev)
S = float(np.float32(np.log2(enc.per_level_scale)))
st = torch.cuda.current_stream().cuda_stream
T = enc.embeddings.shape[0]
B = 5003
g = torch.Generator().manual_seed(0)
o = (torch.rand(B // 50 + 1, 3, generator=g) * 0.4 + 0.5).repeat_interleave(50, dim=0)[:B]
x = (o + (torch.arange(B)[:, None] % 50) * torch.tensor([[3e-4, 2e-4, 1e-4]])).to(dev).contiguous()
for half in (True, False):
    dt = torch.float16 if half else torch.float32
    g0 = torch.randn(B, 32, generator=g).to(dev).to(dt)
    g1 = torch.randn(B, 32, generator=g).to(dev).to(dt)
    ref = None
    for walk in (0, 32, 128):
        lib.nrf_grid_set_bwd_walk(walk)
        gp = torch.zeros(T, 2, 2, device=dev)
        assert lib.nrf_grid_encode_backward_pair(g0.data_ptr(), g1.data_ptr(), x.data_ptr(), enc.offsets.data_ptr(), gp.data_ptr(), B, 16, S, 16,
                                                 0, 1, 0, 1 if half else 0, None, st) == 0
        ge = torch.zeros(T, 2, device=dev)
        assert lib.nrf_grid_encode_backward(g0.data_ptr(), x.data_ptr(), None, enc.offsets.data_ptr(), ge.data_ptr(), B, 3, 2, 16, S, 16, 0, None,
                                            None, 0, 1, 0, 1 if half else 0, 0, 1, st) == 0
        torch.cuda.synchronize()
        if ref is None:
            ref = (gp, ge)
        e = max(float((gp - ref[0]).abs().max() / ref[0].abs().max()), float((ge - ref[1]).abs().max() / ref[1].abs().max()),
                float((gp[:, 0] - ge).abs().max() / ge.abs().max()))
        print('half=%d walk=%3d: rel err vs thread-per-sample / pair vs single %.1e' % (half, walk, e))
        assert e < 1e-5
lib.nrf_grid_set_bwd_walk(128)
print('WALK SANITY OK')
