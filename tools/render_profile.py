"""GPU-time vs wall-time of one full-frame render (config 3) + kernel breakdown: python tools/render_profile.py [graph|host]"""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from torch.profiler import profile, ProfilerActivity
import bench as B
from nerfstyle_b200 import model as M, raymarching, scenes

dev = torch.device('cuda:0')
mode = sys.argv[1] if len(sys.argv) > 1 else 'graph'
torch.manual_seed(0)
m = M.StyleTCNerf([-2., -2., -2.], [2., 2., 2.], class_dim=B.N_CLASSES).to(dev)
r = M.Renderer(m, 2.0, raymarch_channels=3 + B.N_CLASSES, density_scale=50.0).to(dev)
r.density_bitfield = raymarching.packbits(scenes.analytic_density_grid(2, 128, 2.0).to(dev), 0.5)
intr = scenes.scaled_intrinsics(1008, 756)
pose = scenes.synthetic_poses(8, 1)[1]
idx = torch.arange(0, 1008 * 756, device=dev)
o, d = scenes.generate_rays(pose, intr, dev, idx)


def frame():
    with torch.no_grad(), torch.autocast('cuda', dtype=torch.float16):
        return r.render_test_graph(o, d) if mode == 'graph' else r.render_test(o, d, sync_every=4)


for _ in range(2):
    frame()
torch.cuda.synchronize()
t0 = time.perf_counter()
frame()
torch.cuda.synchronize()
print('%s loop: wall %.2f ms' % (mode, (time.perf_counter() - t0) * 1e3))
if mode == 'graph':
    print('iterations:', int(r._gs['ctl'][6]))
with profile(activities=[ProfilerActivity.CPU, ProfilerActivity.CUDA]) as prof:
    frame()
    torch.cuda.synchronize()
ka = [k for k in prof.key_averages() if k.self_device_time_total > 0]
tot = sum(k.self_device_time_total for k in ka)
print('total device time %.2f ms' % (tot / 1e3))
for k in sorted(ka, key=lambda k: -k.self_device_time_total)[:14]:
    print('%8.3f ms %5.1f%% x%-4d %s' % (k.self_device_time_total / 1e3, 100 * k.self_device_time_total / tot, k.count, k.key[:90]))
