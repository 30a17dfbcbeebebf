"""Where does the e2e (host inputs + per-step loss read-back) time go?  Variants of the bench loop on a fresh trainer each."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import bench as B

dev = torch.device('cuda:0')
W, K = 8, 32
host, devb = B.make_batches(W + K, 8192, 0, 1, dev)


def run(name, h2d, item, small=False):
    torch.manual_seed(0); torch.cuda.manual_seed_all(0)
    ts = B.build_trainer(dev, True, 1)
    stage = torch.empty_like(devb[0])
    for s in range(W):
        ts.step(*B.unpack(devb[s]))
    torch.cuda.synchronize()
    cpu_in_step = 0.0
    t0 = time.perf_counter()
    for s in range(W, W + K):
        if h2d:
            stage.copy_(host[s], non_blocking=True)
            src = stage
        else:
            src = devb[s]
        c0 = time.perf_counter()
        loss = ts.step(*B.unpack(src))
        cpu_in_step += time.perf_counter() - c0
        if item:
            float(loss.item())
    torch.cuda.synchronize()
    dt = (time.perf_counter() - t0) / K * 1e3
    print('%-40s %.3f ms/step   (python time inside step(): %.3f ms)' % (name, dt, cpu_in_step / K * 1e3), flush=True)


run('device inputs, no read-back', False, False)
run('device inputs, loss.item() per step', False, True)
run('H2D per step, no read-back', True, False)
run('H2D + loss.item() per step (e2e)', True, True)
run('device inputs, no read-back (again)', False, False)
