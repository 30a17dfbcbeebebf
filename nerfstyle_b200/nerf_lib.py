"""Ray generation on the device, behind the reference's `nerf_lib.generate_rays` surface (SURVEY.md 8f NEXT-1).

Reference: NerfLib.generate_rays (nerf_lib.py:69-142), RayBatch (common.py:126-150), Intrinsics (common.py:41-114),
Box2D (common.py:25-38).  Same argument names, meaning, asserts and return shape: `(RayBatch, target)`.

What changes underneath: the reference builds the whole frame's pixel grid with numpy on the host every call,
uploads it, rotates every pixel and only then picks `bsize` of them with `np.random.choice(..., replace=False)` (an
O(W*H) host permutation); here ONE kernel (`nrf_generate_rays`, csrc/rays.cu) computes just the K selected rays --
directions, unit normalisation, tiled origins and the target-pixel gather -- and the without-replacement draw is a
device `torch.randperm`.  The draw is therefore a different (equally uniform) random stream than numpy's; pass
`indices=` to choose the pixels yourself (that is how the parity tests replay the reference's draw).
"""
from dataclasses import dataclass
from typing import Optional

import numpy as np
import torch
import torch.nn.functional as F

from . import _lib as L


@dataclass(frozen=True)
class Box2D:
    """common.py:25-38"""
    x: int
    y: int
    w: int
    h: int

    def wrange(self):
        return slice(self.x, self.x + self.w)

    def hrange(self):
        return slice(self.y, self.y + self.h)


@dataclass(frozen=True)
class Intrinsics:
    """common.py:41-114 (field order h, w, fx, fy, cx, cy)."""
    h: int
    w: int
    fx: float
    fy: float
    cx: float
    cy: float

    def __post_init__(self):
        object.__setattr__(self, 'h', int(self.h))
        object.__setattr__(self, 'w', int(self.w))

    def size(self):
        return self.w, self.h

    def scale(self, w, h):
        cx, cy = w / 2., h / 2.
        old_ar, new_ar = self.w / self.h, w / h
        ratio = h / self.h if new_ar >= old_ar else w / self.w
        return Intrinsics(h, w, self.fx * ratio, self.fy * ratio, cx, cy)


@dataclass
class RayBatch:
    """common.py:126-150.  Built by generate_rays with the directions already unit length (the kernel normalises);
    constructing one by hand normalises like the reference."""
    origins: torch.Tensor
    dirs: torch.Tensor
    _normalized: bool = False

    def __post_init__(self):
        assert len(self.origins.shape) <= 2
        assert len(self.dirs.shape) == 2
        if len(self.origins.shape) == 1:
            self.origins = torch.tile(self.origins, (len(self.dirs), 1))
        assert self.origins.shape == self.dirs.shape
        if not self._normalized:
            self.dirs = self.dirs / torch.norm(self.dirs, dim=-1, keepdim=True)

    def __len__(self):
        return len(self.dirs)


class NerfLib:
    """nerf_lib.py:23-142: the module-level object the reference's trainers call (`nerf_lib.device = ...;
    nerf_lib.generate_rays(...)`)."""

    def __init__(self):
        self._device = None
        self._ready = False

    @property
    def device(self):
        return self._device

    @device.setter
    def device(self, device):
        device = torch.device(device)
        assert device.type == 'cuda', 'Device must be GPU'
        self._device = device
        self._ready = True

    def generate_rays(self, pose, intr, img=None, patch: Optional[Box2D] = None, precrop: float = 1.,
                      bsize: Optional[int] = None, camera_flip: int = 0, indices=None, generator=None):
        """Returns (RayBatch of K rays, target [K,3] or None).  `indices` / `generator` are additions: the flat
        window-pixel ids to use instead of a fresh draw, and the torch generator of that draw."""
        assert self._ready, 'Please assign a GPU to nerf_lib.device first.'
        assert (precrop >= 0.) and (precrop <= 1.)
        assert (precrop >= 1.) or (patch is None), 'Using both precrop and patch is not supported'
        dev = self._device
        fw, fh = intr.size()
        w, h, dx, dy = intr.w, intr.h, 0, 0
        if precrop < 1.:                                    # nerf_lib.py:109-112
            w, h = int(intr.w * precrop), int(intr.h * precrop)
            dx, dy = (intr.w - w) // 2, (intr.h - h) // 2
        x0, y0, win_w, win_h = dx, dy, w, h
        if patch is not None:                               # :114-116 (slices clip like numpy's)
            px0, px1 = min(patch.x, win_w), min(patch.x + patch.w, win_w)
            py0, py1 = min(patch.y, win_h), min(patch.y + patch.h, win_h)
            x0, y0, win_w, win_h = x0 + px0, y0 + py0, max(px1 - px0, 0), max(py1 - py0, 0)
        pose = torch.as_tensor(pose, dtype=torch.float32, device=dev)
        if pose.shape != (4, 4):                            # the reference slices [:3,:3] / [:3,3]; accept 3x4 too
            full = torch.eye(4, dtype=torch.float32, device=dev)
            full[:pose.shape[0], :pose.shape[1]] = pose
            pose = full
        pose = pose.contiguous()
        target = None
        gather = None
        if bsize is None:
            idx, K = None, win_w * win_h
            if img is not None:                             # :128-131: the WHOLE image, also under precrop
                if fh != img.shape[-2] or fw != img.shape[-1]:
                    img = F.interpolate(img.unsqueeze(0), size=(fh, fw)).squeeze(0)
                target = img.permute(1, 2, 0).reshape(-1, img.shape[0])
        else:
            if patch is not None:
                # the reference draws over the un-patched w*h and indexes the patch-sized grid with it (:132-135)
                raise IndexError('generate_rays: bsize together with patch indexes outside the patch in the reference')
            if indices is None:                             # :132 np.random.choice(arange(w*h), bsize, replace=False)
                if bsize > w * h:
                    raise ValueError("Cannot take a larger sample than population when 'replace=False'")
                indices = torch.randperm(w * h, device=dev, generator=generator)[:bsize]
            idx = torch.as_tensor(indices, device=dev).to(torch.int64).contiguous()
            K = idx.numel()
            if img is not None:
                gather = torch.as_tensor(img, device=dev)
                if gather.dtype != torch.float32 or gather.shape[0] != 3:
                    # uncommon layouts go through torch indexing, like the reference (:135-137)
                    target = gather.permute(1, 2, 0)[idx // w + dy, idx % w + dx]
                    gather = None
                else:
                    gather = gather.contiguous()
                    target = torch.empty(K, 3, dtype=torch.float32, device=dev)
        rays_o = torch.empty(K, 3, dtype=torch.float32, device=dev)
        rays_d = torch.empty(K, 3, dtype=torch.float32, device=dev)
        f32 = np.float32
        with torch.cuda.device(dev):
            L.check(L.lib().nrf_generate_rays(
                pose.data_ptr(), float(f32(intr.fx)), float(f32(intr.fy)), float(f32(intr.cx)), float(f32(intr.cy)),
                int(x0), int(y0), int(max(win_w, 1)), int(K), L.ptr(idx), int(camera_flip), L.ptr(gather),
                int(gather.shape[2]) if gather is not None else 0, int(gather.shape[1]) if gather is not None else 0,
                rays_o.data_ptr(), rays_d.data_ptr(), L.ptr(target) if gather is not None else None,
                L.stream_of(rays_o)), 'generate_rays')
        return RayBatch(rays_o, rays_d, _normalized=True), target


nerf_lib = NerfLib()
