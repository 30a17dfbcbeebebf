"""Synthetic, seeded inputs shaped like the reference's LLFF "room" set-up (BASELINE.json configs 1-3).

Nothing here reads /root/reference: the camera rig is synthesised (forward-facing, LLFF-like: 504x378,
fl 383.83, translations scaled by 0.33, camera_flip=3) and so are the target pixels / labels.  Ray
generation follows the arithmetic of nerf_lib.generate_rays (nerf_lib.py:69-142) and RayBatch
normalisation (common.py:139-147), but runs on the device (SURVEY.md 8f NEXT-1).
"""
import math

import numpy as np
import torch

ROOM = dict(w=504, h=378, fx=383.829783860205, fy=383.829783860205, cx=252.0, cy=189.0, bound=2.0, scale=0.33,
            flip_camera=3, n_train=35)


def scaled_intrinsics(w, h, base=ROOM):
    """Intrinsics.scale (common.py:92-114)."""
    old_ar = base['w'] / base['h']
    new_ar = w / h
    ratio = h / base['h'] if new_ar >= old_ar else w / base['w']
    return dict(base, w=w, h=h, fx=base['fx'] * ratio, fy=base['fy'] * ratio, cx=w / 2., cy=h / 2.)


def synthetic_poses(n=35, seed=0, scale=ROOM['scale']):
    """Forward-facing camera-to-world matrices: cameras on a small patch at x ~ 3.9 (unscaled) looking towards -x
    (in the flipped camera convention), like an LLFF capture.  Returns float32 [n,4,4]; translations * scale."""
    rng = np.random.RandomState(seed)
    poses = np.zeros((n, 4, 4), np.float32)
    for i in range(n):
        eye = np.array([3.9 + 0.05 * rng.randn(), 0.6 * (rng.rand() - 0.5), 0.45 * (rng.rand() - 0.5) + 0.3])
        target = np.array([0.0, 0.15 * rng.randn(), 0.15 * rng.randn()])
        fwd = target - eye
        fwd /= np.linalg.norm(fwd)
        up = np.array([0.0, 0.0, 1.0])
        right = np.cross(fwd, up)
        right /= np.linalg.norm(right)
        true_up = np.cross(right, fwd)
        # columns: camera x (right), camera y (up), camera z (backwards); flip_camera=3 negates y and z of the
        # pixel directions (nerf_lib.py:121-122), so a pixel at the image centre looks along -z_cam = fwd
        R = np.stack([right, true_up, -fwd], axis=1)
        poses[i, :3, :3] = R
        poses[i, :3, 3] = eye * scale
        poses[i, 3, 3] = 1.0
    return poses


def generate_rays(pose, intr, device, indices=None):
    """Pixel-centre rays of one view.  pose [4,4]; indices: optional LongTensor of flat pixel ids (row-major over
    h x w).  Returns unit-norm rays_d [K,3] and rays_o [K,3] on `device` (float32)."""
    w, h = int(intr['w']), int(intr['h'])
    if torch.device(device).type == 'cuda':
        # the product path: one nrf_generate_rays launch (nerfstyle_b200.nerf_lib), same arithmetic as the lines below
        from .nerf_lib import Intrinsics, NerfLib
        lib = NerfLib()
        lib.device = device
        it = Intrinsics(h, w, intr['fx'], intr['fy'], intr['cx'], intr['cy'])
        if indices is None:
            rb, _ = lib.generate_rays(pose, it, camera_flip=int(intr.get('flip_camera', 0)))
        else:
            rb, _ = lib.generate_rays(pose, it, bsize=int(indices.numel()), indices=indices,
                                      camera_flip=int(intr.get('flip_camera', 0)))
        return rb.origins, rb.dirs
    pose = torch.as_tensor(pose, dtype=torch.float32, device=device)
    if indices is None:
        indices = torch.arange(w * h, device=device)
    iy = torch.div(indices, w, rounding_mode='floor')
    ix = indices - iy * w
    # np.linspace(0, w, 2w+1)[1::2] = 0.5, 1.5, ...
    px = ix.to(torch.float32) + 0.5
    py = iy.to(torch.float32) + 0.5
    dirs = torch.stack([(px - intr['cx']) / intr['fx'], (py - intr['cy']) / intr['fy'], torch.ones_like(px)], dim=-1)
    flip = torch.tensor([-1.0 if (intr.get('flip_camera', 0) >> s) & 1 else 1.0 for s in (2, 1, 0)], device=device)
    dirs = dirs * flip
    rays_d = dirs @ pose[:3, :3].T
    rays_d = rays_d / torch.norm(rays_d, dim=-1, keepdim=True)
    rays_o = pose[:3, 3].expand_as(rays_d).contiguous()
    return rays_o, rays_d


def synthetic_target(indices, intr, n_classes=8):
    """Smooth synthetic RGB target + segmentation label per pixel."""
    w = int(intr['w'])
    iy = torch.div(indices, w, rounding_mode='floor').to(torch.float32)
    ix = (indices % w).to(torch.float32)
    rgb = torch.stack([0.5 + 0.5 * torch.sin(ix * 0.05), 0.5 + 0.5 * torch.sin(iy * 0.07),
                       0.5 + 0.5 * torch.sin((ix + iy) * 0.03)], dim=-1)
    seg = ((ix // max(w // n_classes, 1)).long() + 2 * (iy // max(int(intr['h']) // 2, 1)).long()) % n_classes
    return rgb, seg


def random_rays(n, seed=0, device='cpu'):
    """Config-1 style rays: origins U(-0.5,0.5)^3, directions normalised N(0,1)^3 (CPU generator, moved)."""
    g = torch.Generator().manual_seed(seed)
    o = torch.rand(n, 3, generator=g) - 0.5
    d = torch.randn(n, 3, generator=g)
    d = d / d.norm(dim=-1, keepdim=True)
    return o.to(device), d.to(device)


def analytic_density_grid(cascade=2, H=128, bound=2.0):
    """Occupancy (a) of SURVEY.md 8d: a cell is occupied iff its centre has max-norm < 1.5 and L2 norm > 0.3.
    Returns a float density grid [cascade, H^3] in Morton order (1 = occupied, 0 = empty), CPU tensor."""
    idx = torch.arange(H ** 3, dtype=torch.int64)

    def compact(v):
        v = v & 0x49249249
        v = (v | (v >> 2)) & 0xc30c30c3
        v = (v | (v >> 4)) & 0x0f00f00f
        v = (v | (v >> 8)) & 0xff0000ff
        v = (v | (v >> 16)) & 0x0000ffff
        return v
    x, y, z = compact(idx), compact(idx >> 1), compact(idx >> 2)
    grid = torch.zeros(cascade, H ** 3)
    for cas in range(cascade):
        b = min(2.0 ** cas, bound)
        c = torch.stack([x, y, z], dim=-1).to(torch.float32)
        centre = ((c + 0.5) / H * 2 - 1) * b
        occ = (centre.abs().amax(dim=-1) < 1.5) & (centre.norm(dim=-1) > 0.3)
        grid[cas] = occ.to(torch.float32)
    return grid


def bernoulli_density_grid(cascade=2, H=128, p=0.5, seed=1):
    g = torch.Generator().manual_seed(seed)
    return (torch.rand(cascade, H ** 3, generator=g) < p).to(torch.float32)


def frame_indices(intr, n_rays, generator):
    """A random subset of pixel ids without replacement (nerf_lib.py:134 uses np.random.choice)."""
    total = int(intr['w']) * int(intr['h'])
    if n_rays <= total:
        return torch.randperm(total, generator=generator)[:n_rays]
    # more rays than pixels (ray-batch sweeps): whole extra permutations of the frame, so the batch really has n_rays rays
    reps = -(-n_rays // total)
    return torch.cat([torch.randperm(total, generator=generator) for _ in range(reps)])[:n_rays]


def psnr(mse):
    return -10.0 * math.log(max(float(mse), 1e-20)) / math.log(10.0)
