"""Does it train?  400 steps of the bench workload; prints the loss / PSNR curve and the sample count per step."""
import os, sys, math
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import bench as B

dev = torch.device('cuda:0')
n_steps = int(sys.argv[1]) if len(sys.argv) > 1 else 400
host, devb = B.make_batches(n_steps, 8192, 0, 1, dev)
torch.manual_seed(0)
ts = B.build_trainer(dev, True, 1)
for s in range(n_steps):
    loss = ts.step(*B.unpack(devb[s]))
    if s % 50 == 0 or s == n_steps - 1:
        r = ts.renderer
        lv = float(loss)
        print('step %4d  loss %.5f  psnr(all terms) %.2f dB  samples %d  scale %.0f  good_steps %d' % (
            s, lv, -10 * math.log10(max(lv, 1e-12)), int(r.step_counter[(r.local_step - 1) % 16, 0]), float(ts.fused.scale),
            int(ts.fused.good_steps)), flush=True)
assert math.isfinite(lv)          # the synthetic targets are not multi-view consistent: the loss plateaus, it must stay finite
print('TRAIN SANITY OK')
