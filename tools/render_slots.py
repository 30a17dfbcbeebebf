"""How many of the sample slots a full-frame render processes are real samples?  (config 3, trained-like case)
Replays the reference loop (renderer.py:249-286) and counts, per iteration, n_alive * n_step slots vs slots with delta > 0."""
import os, sys
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
N = o.shape[0]
with torch.no_grad(), torch.autocast('cuda', dtype=torch.float16):
    nears, fars = raymarching.near_far_from_aabb(o, d, r.aabb, r.min_near)
    ws = torch.zeros(N, device=dev); depth = torch.zeros(N, device=dev); image = torch.zeros(N, 11, device=dev)
    alive = torch.arange(N, dtype=torch.int32, device=dev)
    rays_t = nears.clone()[:, None]
    step, it, tot_slots, tot_real = 0, 0, 0, 0
    while step < r.max_steps:
        n_alive = len(alive)
        if n_alive <= 0:
            break
        n_step = max(min(N // n_alive, 8), 1)
        xyzs, dirs, deltas = raymarching.march_rays(n_alive, n_step, alive, rays_t, o, d, None, r.bound, r.density_bitfield,
                                                    r.cascade, r.grid_size, nears, fars, 128, False, 0., r.max_steps, False)
        real = int((deltas[:n_alive * n_step, 0] > 0).sum())
        rgbs, sigmas = r.model(xyzs, dirs=dirs)
        sigmas = sigmas * r.density_scale
        raymarching.composite_rays(n_alive, n_step, alive, rays_t, sigmas, rgbs, deltas, False, ws, depth, image, r.t_thresh)
        alive = alive[alive >= 0]
        if it < 12 or it % 8 == 0:
            print('it %3d  n_alive %7d  n_step %d  slots %8d  real %8d (%.0f%%)' % (it, n_alive, n_step, n_alive * n_step, real,
                                                                                 100.0 * real / (n_alive * n_step)))
        tot_slots += n_alive * n_step; tot_real += real
        step += n_step; it += 1
print('iterations %d, slots %d, real samples %d (%.1f%%), %.1f real samples / ray' % (it, tot_slots, tot_real, 100.0 * tot_real / tot_slots, tot_real / N))
