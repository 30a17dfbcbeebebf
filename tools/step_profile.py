"""torch.profiler view of a few bench.py train steps (kernel-time table + CPU-side op table)."""
import os
import sys

import torch
from torch.profiler import ProfilerActivity, profile

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402

dev = torch.device('cuda:0')
torch.cuda.set_device(0)
fused = '--torch-optim' not in sys.argv
ts = bench.build_trainer(dev, True, 1)
if not fused:
    from nerfstyle_b200.trainer import TrainStep
    ts = TrainStep(ts.renderer, enable_amp=True, world_size=1, fused_optimizer=False)
host, devb = bench.make_batches(12, 8192, 0, 1, dev)
for s in range(6):
    ts.step(*bench.unpack(devb[s]))
torch.cuda.synchronize()
with profile(activities=[ProfilerActivity.CPU, ProfilerActivity.CUDA]) as prof:
    for s in range(6, 10):
        ts.step(*bench.unpack(devb[s]))
    torch.cuda.synchronize()
print(prof.key_averages().table(sort_by='cuda_time_total', row_limit=45, max_name_column_width=70))
print(prof.key_averages().table(sort_by='self_cpu_time_total', row_limit=25, max_name_column_width=70))
