"""Generates tests/golden/rays.npz by running the REFERENCE's own `NerfLib.generate_rays` (nerf_lib.py:69-142) and
`RayBatch.__post_init__` (common.py:139-147), unmodified, on the CPU of this container.

The reference's `utils` package needs matplotlib / imageio (absent here) but neither function touches it, so an empty
stand-in module is registered under that name before `common` / `nerf_lib` are imported from /root/reference.  The
library object's device is set to CPU directly (its setter insists on a GPU; the function body only uses it as the
`device=` of two tensor constructors).  Unlike the other fixtures these are outputs of the reference itself:
ray-generation parity is PINNED by the reference.  Run from the repo root:  python tests/golden/make_rays_golden.py
"""
import os
import sys
import types

import numpy as np
import torch

REF = '/root/reference'
HERE = os.path.dirname(os.path.abspath(__file__))


def load_reference():
    sys.modules.setdefault('utils', types.ModuleType('utils'))
    sys.path.insert(0, REF)
    import common       # noqa: E402
    import nerf_lib     # noqa: E402
    lib = nerf_lib.NerfLib()
    lib._device = torch.device('cpu')
    lib._ready = True
    return common, lib


def pose_of(seed):
    rs = np.random.RandomState(seed)
    q, _ = np.linalg.qr(rs.randn(3, 3))
    pose = np.eye(4, dtype=np.float32)
    pose[:3, :3] = q.astype(np.float32)
    pose[:3, 3] = (rs.randn(3) * 0.7).astype(np.float32)
    return pose


def synth_img(w, h):
    """Closed-form image [3,h,w] (the tests rebuild it instead of storing it)."""
    c, y, x = np.meshgrid(np.arange(3), np.arange(h), np.arange(w), indexing='ij')
    return (((x * 7 + y * 13 + c * 29) % 251) / 251.0).astype(np.float32)


def main():
    common, lib = load_reference()
    base = common.Intrinsics(378, 504, 383.829783860205, 383.829783860205, 252.0, 189.0)
    out = {}
    cases = [
        # name, (w, h), camera_flip, precrop, patch (x, y, w, h), bsize, seed
        ('full_small', (24, 18), 3, 1.0, None, None, 1),
        ('full_noflip', (21, 13), 0, 1.0, None, None, 2),
        ('flip5', (16, 12), 5, 1.0, None, None, 3),
        ('precrop', (40, 30), 3, 0.5, None, None, 4),
        ('patch', (40, 30), 3, 1.0, (7, 5, 12, 9), None, 5),
        ('sampled', (504, 378), 3, 1.0, None, 2048, 6),
        ('sampled_precrop', (504, 378), 3, 0.5, None, 1024, 7),
        ('native', (504, 378), 3, 1.0, None, None, 8),
    ]
    names = []
    for name, (w, h), flip, precrop, patch, bsize, seed in cases:
        intr = base.scale(w, h) if (w, h) != (504, 378) else base
        pose = pose_of(seed)
        img = torch.from_numpy(synth_img(intr.w, intr.h))
        box = common.Box2D(*patch) if patch is not None else None
        np.random.seed(seed)                       # np.random.choice inside generate_rays (nerf_lib.py:132)
        rays, target = lib.generate_rays(torch.from_numpy(pose), intr, img=None if patch is not None else img, patch=box,
                                         precrop=precrop, bsize=bsize, camera_flip=flip)
        if bsize is not None:                      # replay the draw to record which pixels were taken
            np.random.seed(seed)
            cw, ch = (int(intr.w * precrop), int(intr.h * precrop)) if precrop < 1. else (intr.w, intr.h)
            idx = np.random.choice(np.arange(cw * ch), bsize, replace=False)
        else:
            idx = np.zeros(0, np.int64)
        out[name + '/pose'] = pose
        out[name + '/intr'] = np.array([intr.w, intr.h, intr.fx, intr.fy, intr.cx, intr.cy], np.float64)
        out[name + '/args'] = np.array([flip, precrop, -1 if bsize is None else bsize] + list(patch or (-1, -1, -1, -1)), np.float64)
        out[name + '/indices'] = idx.astype(np.int64)
        out[name + '/origins'] = rays.origins.numpy()
        out[name + '/dirs'] = rays.dirs.numpy()
        keep_t = target is not None and name != 'native'
        out[name + '/target'] = target.numpy() if keep_t else np.zeros(0, np.float32)
        names.append(name)
        print(name, rays.dirs.shape, None if target is None else tuple(target.shape))
    # keep the committed file small: the two big full-resolution cases store every 97th ray only
    for name in ('native',):
        sel = np.arange(0, out[name + '/dirs'].shape[0], 97)
        out[name + '/indices'] = sel.astype(np.int64)          # flat pixel ids of the stored rows
        out[name + '/origins'] = out[name + '/origins'][sel]
        out[name + '/dirs'] = out[name + '/dirs'][sel]
        out[name + '/args'][2] = -2                            # marker: full frame, rows sub-sampled for storage
    out['names'] = np.array(names)
    np.savez_compressed(os.path.join(HERE, 'rays.npz'), **out)
    print('wrote rays.npz', os.path.getsize(os.path.join(HERE, 'rays.npz')))


if __name__ == '__main__':
    main()
