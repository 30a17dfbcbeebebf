"""GPU-time vs wall-time of one full-frame render (config 3) + kernel breakdown: python tools/render_profile.py [sync_every]"""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from torch.profiler import profile, ProfilerActivity
import bench as B
from nerfstyle_b200 import model as M, raymarching, scenes

dev = torch.device('cuda:0')
se = int(sys.argv[1]) if len(sys.argv) > 1 else 4
torch.manual_seed(0)
m = M.StyleTCNerf([-2., -2., -2.], [2., 2., 2.], class_dim=B.N_CLASSES).to(dev)
r = M.Renderer(m, 2.0, raymarch_channels=3 + B.N_CLASSES, density_scale=50.0).to(dev)
r.density_bitfield = raymarching.packbits(scenes.analytic_density_grid(2, 128, 2.0).to(dev), 0.5)
intr = scenes.scaled_intrinsics(1008, 756)
pose = scenes.synthetic_poses(8, 1)[1]
idx = torch.arange(0, 1008 * 756, device=dev)
o, d = scenes.generate_rays(pose, intr, dev, idx)
for _ in range(2):
    with torch.no_grad(), torch.autocast('cuda', dtype=torch.float16):
        r.render_test(o, d, sync_every=se)
torch.cuda.synchronize()
t0 = time.perf_counter()
with torch.no_grad(), torch.autocast('cuda', dtype=torch.float16):
    r.render_test(o, d, sync_every=se)
torch.cuda.synchronize()
print('wall %.2f ms' % ((time.perf_counter() - t0) * 1e3))
with profile(activities=[ProfilerActivity.CPU, ProfilerActivity.CUDA]) as prof:
    with torch.no_grad(), torch.autocast('cuda', dtype=torch.float16):
        r.render_test(o, d, sync_every=se)
    torch.cuda.synchronize()
ka = prof.key_averages()
tot = sum(k.self_device_time_total for k in ka)
print('total device time %.2f ms' % (tot / 1e3))
print(ka.table(sort_by='self_cuda_time_total', row_limit=22, max_name_column_width=70))
