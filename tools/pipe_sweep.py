"""Train-step time against the number of chunks of the optional gather -> networks stream pipeline (gridencoder.PIPELINE_CHUNKS;
1 = off, the default).  Usage (GPU box): python tools/pipe_sweep.py"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import bench as B
from nerfstyle_b200 import gridencoder as GE
dev = torch.device('cuda:0')
from nerfstyle_b200.trainer import TrainStep
TrainStep.reserve_workspace(dev)
host, devb = B.make_batches(40, 8192, 0, 1, dev)
for n, minrows in ((1, 1 << 19), (2, 1 << 19), (4, 1 << 19), (8, 1 << 18), (16, 1 << 17)):
    GE.PIPELINE_CHUNKS[0] = n
    GE._pipeline_chunks.__defaults__ = (minrows, None)
    torch.manual_seed(0); torch.cuda.manual_seed_all(0)
    ts = B.build_trainer(dev, True, 1)
    for s in range(8):
        ts.step(*B.unpack(devb[s]))
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for s in range(8, 40):
        ts.step(*B.unpack(devb[s]))
    e1.record(); torch.cuda.synchronize()
    print('chunks %2d: %.4f ms/step' % (n, e0.elapsed_time(e1) / 32))
    del ts
