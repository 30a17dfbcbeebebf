"""Frame time of the device-driven render loop against steps_per_iteration (config 3, trained-like case)."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import bench as B
from nerfstyle_b200 import model as M, raymarching, scenes

dev = torch.device('cuda:0')
torch.manual_seed(0)
m = M.StyleTCNerf([-2., -2., -2.], [2., 2., 2.], class_dim=B.N_CLASSES).to(dev)
r = M.Renderer(m, 2.0, raymarch_channels=3 + B.N_CLASSES, density_scale=50.0).to(dev)
r.density_bitfield = raymarching.packbits(scenes.analytic_density_grid(2, 128, 2.0).to(dev), 0.5)
intr = scenes.scaled_intrinsics(1008, 756)
pose = scenes.synthetic_poses(8, 1)[1]
o, d = scenes.generate_rays(pose, intr, dev, torch.arange(0, 1008 * 756, device=dev))
for spi, ce in ((4, 4), (8, 4), (8, 2), (8, 1), (8, 8)):
    ts = []
    for f in range(4):
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        with torch.no_grad(), torch.autocast('cuda', dtype=torch.float16):
            img, _, _ = r.render_test_graph(o, d, steps_per_iteration=spi, check_every=ce)
        float(img.sum().item())
        if f > 0:
            ts.append((time.perf_counter() - t0) * 1e3)
    print('steps_per_iteration %d check_every %d: %.2f ms/frame, %d iterations' % (spi, ce, sum(ts) / len(ts), int(r._gs['ctl'][6])))
