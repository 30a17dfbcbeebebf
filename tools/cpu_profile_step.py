"""Host-side (Python) cost of one train step: cProfile over 24 steps with a device sync per step."""
import cProfile, os, pstats, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import bench as B

dev = torch.device('cuda:0')
host, devb = B.make_batches(32, 8192, 0, 1, dev)
ts = B.build_trainer(dev, True, 1)
for s in range(8):
    ts.step(*B.unpack(devb[s]))
torch.cuda.synchronize()
pr = cProfile.Profile()
pr.enable()
for s in range(8, 32):
    ts.step(*B.unpack(devb[s]))
    torch.cuda.synchronize()
pr.disable()
st = pstats.Stats(pr)
st.sort_stats('tottime').print_stats(28)
