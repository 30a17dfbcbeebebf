"""CPU oracle for ray generation -- TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).

Restates NerfLib.generate_rays (/root/reference/nerf_lib.py:69-142) and RayBatch.__post_init__ (common.py:139-147)
in numpy float32, operation for operation.  Pinned by tests/golden/rays.npz, which holds outputs of the reference's own
function run on the CPU of the build container (tests/golden/make_rays_golden.py): parity pinned by the reference.
"""
import numpy as np


def generate_rays(pose, w, h, fx, fy, cx, cy, img=None, patch=None, precrop=1., indices=None, camera_flip=0):
    """pose [4,4] f32; patch = (x, y, w, h) or None; indices = flat ids over the (pre)cropped window or None (all rays).
    Returns origins [K,3], dirs [K,3] (unit), target [K,3] or None."""
    pose = np.asarray(pose, np.float32)
    fw, fh = w, h
    x_coords = np.linspace(0, fw, num=2 * fw + 1, dtype=np.float32)[1::2]          # nerf_lib.py:103-104
    y_coords = np.linspace(0, fh, num=2 * fh + 1, dtype=np.float32)[1::2]
    dx = dy = 0
    pose_r, pose_t = pose[:3, :3], pose[:3, 3]
    if precrop < 1.:                                                                # :109-112
        w, h = int(fw * precrop), int(fh * precrop)
        dx, dy = (fw - w) // 2, (fh - h) // 2
        x_coords, y_coords = x_coords[dx:dx + w], y_coords[dy:dy + h]
    if patch is not None:                                                           # :114-116
        x_coords = x_coords[patch[0]:patch[0] + patch[2]]
        y_coords = y_coords[patch[1]:patch[1] + patch[3]]
    i, j = np.meshgrid(x_coords, y_coords, indexing='xy')                           # :118
    k = np.ones_like(i)
    dirs = np.stack([(i - np.float32(cx)) / np.float32(fx), (j - np.float32(cy)) / np.float32(fy), k], axis=-1)
    flip = np.where([(camera_flip >> s) & 1 for s in [2, 1, 0]], -1, 1).astype(np.float32)   # :122-123
    dirs = (dirs * flip).astype(np.float32)
    # einsum('ij, hwj -> hwi') in float32, j accumulated in order with fused multiply-adds (what the BLAS kernels and
    # the device kernel do); emulated exactly through float64 products of float32 values rounded once per step
    rays_d = np.zeros(dirs.shape, np.float32)
    for a in range(3):
        acc = (pose_r[a, 0] * dirs[..., 0]).astype(np.float32)
        for b in (1, 2):
            acc = (np.float64(pose_r[a, b]) * dirs[..., b].astype(np.float64) + acc.astype(np.float64)).astype(np.float32)
        rays_d[..., a] = acc
    target = None
    if indices is None:                                                             # :128-131
        rays_d = rays_d.reshape(-1, 3)
        if img is not None:
            target = np.transpose(img, (1, 2, 0)).reshape(-1, img.shape[0])
    else:                                                                           # :132-137
        indices = np.asarray(indices, np.int64)
        r, c = indices // w, indices % w
        rays_d = rays_d[r, c]
        if img is not None:
            target = np.transpose(img, (1, 2, 0))[r + dy, c + dx]
    n = np.sqrt((rays_d.astype(np.float64) ** 2).sum(-1)).astype(np.float32)        # common.py:147
    dirs_n = (rays_d / n[:, None]).astype(np.float32)
    origins = np.tile(pose_t, (len(dirs_n), 1)).astype(np.float32)                  # common.py:143-144
    return origins, dirs_n, target
