"""Why is bench.py's first timed phase sometimes slower than the second?  Runs the device-resident loop several times in
one process (fresh trainer each) and prints ms/step, cudaMalloc counts and the slowest steps of each repeat."""
import os
import sys
import time

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench as B  # noqa: E402

dev = torch.device('cuda:0')
W, K = 8, 32
host, devb = B.make_batches(W + K, 8192, 0, 1, dev)
for rep in range(4):
    torch.manual_seed(0)
    torch.cuda.manual_seed_all(0)
    ts = B.build_trainer(dev, True, 1)
    for s in range(W):
        ts.step(*B.unpack(devb[s]))
    torch.cuda.synchronize()
    seg0 = torch.cuda.memory_stats(dev).get('segment.all.allocated', 0)
    evs = [torch.cuda.Event(enable_timing=True) for _ in range(K + 1)]
    host_t = []
    evs[0].record()
    for s in range(W, W + K):
        t0 = time.perf_counter()
        ts.step(*B.unpack(devb[s]))
        host_t.append((time.perf_counter() - t0) * 1e3)
        evs[s - W + 1].record()
    torch.cuda.synchronize()
    seg1 = torch.cuda.memory_stats(dev).get('segment.all.allocated', 0)
    per = [evs[i].elapsed_time(evs[i + 1]) for i in range(K)]
    print('rep %d: %.3f ms/step  (median %.3f, max %.3f at step %d)  host ms/step median %.3f max %.3f  cudaMallocs %d  reserved %.1f GB'
          % (rep, sum(per) / K, sorted(per)[K // 2], max(per), per.index(max(per)) + W, sorted(host_t)[K // 2], max(host_t),
             seg1 - seg0, torch.cuda.memory_reserved(dev) / 2 ** 30), flush=True)
    print('   per-step GPU ms:', ' '.join('%.1f' % v for v in per))
    print('   per-step host ms:', ' '.join('%.1f' % v for v in host_t))
    del ts
