"""Baseline arm: the reference's OWN CUDA extensions (rebuilt for sm_100a by oracle/build_ref.sh into oracle/_ref/,
binaries only) driven with the host-side behaviour of the reference's Python wrappers, on the same workload as
bench.py.  Measurement infrastructure only -- nothing here is on the product path.

What is reproduced from the reference's host side (because it is part of what its step costs):
  * raymarching.py:228-283  zero-filled [N*max_steps, 3|3|4] outputs, `.item()` read of the sample count, slicing to the
    128-aligned count, torch.cuda.empty_cache() every call;
  * raymarching.py:339-345  zero-filled grad_sigmas / grad_rgbs / rgbs_buf in the compositing backward;
  * grid.py:46,58,80-82     level-major [L,B,C] outputs + permute/reshape copy, permuted contiguous grads, zeros_like table
    grads in the (half) table dtype, half tables under autocast;
  * style_nerf.py:144-159   chunks of 10^6 points;
  * trainers/base.py:216-229,405-426  autocast + GradScaler + torch.optim.Adam(eps=1e-15) + EMA.
tiny-cuda-nn is not available offline, so its FullyFusedMLPs are replaced by the stand-in BASELINE.md names: a torch fp16
F.linear chain (cuBLAS) with tcnn's padding (in/out to 16, batch to 128), fp16 output.
"""
import importlib.util
import os

import numpy as np
import torch
import torch.nn as nn
import torch.nn.functional as F
from torch.autograd import Function
from torch.amp import custom_bwd, custom_fwd

ROOT = os.path.dirname(os.path.abspath(__file__))
REFDIR = os.path.join(ROOT, 'oracle', '_ref')


def _load(name):
    path = os.path.join(REFDIR, name + '.so')
    if not os.path.exists(path):
        raise FileNotFoundError(path)
    spec = importlib.util.spec_from_file_location(name, path)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


class _RefOps:
    rm = None
    ge = None


def ref():
    if _RefOps.rm is None:
        _RefOps.rm = _load('_raymarching_ref')
        _RefOps.ge = _load('_gridencoder_ref')
    return _RefOps


def ref_near_far(o, d, aabb, min_near):
    N = o.shape[0]
    nears = torch.empty(N, device=o.device)
    fars = torch.empty(N, device=o.device)
    ref().rm.near_far_from_aabb(o, d, aabb, N, min_near, nears, fars)
    return nears, fars


def ref_march_rays_train(o, d, bound, bitfield, C, H, nears, fars, counter, align=128, max_steps=1024):
    N = o.shape[0]
    M = N * max_steps
    dev = o.device
    xyzs = torch.zeros(M, 3, device=dev)
    dirs = torch.zeros(M, 3, device=dev)
    deltas = torch.zeros(M, 4, device=dev)
    rays = torch.empty(N, 3, dtype=torch.int32, device=dev)
    noises = torch.zeros(N, device=dev)
    ref().rm.march_rays_train(o, d, torch.tensor((), device=dev), bitfield, bound, 0.0, max_steps, False, N, C, H, M, nears, fars,
                              xyzs, dirs, deltas, rays, counter, noises)
    m = counter[0].item()
    m += align - m % align
    xyzs, dirs, deltas = xyzs[:m], dirs[:m], deltas[:m]
    torch.cuda.empty_cache()
    return xyzs, dirs, deltas, rays


class RefComposite(Function):
    @staticmethod
    @custom_fwd(device_type='cuda', cast_inputs=torch.float32)
    def forward(ctx, sigmas, rgbs, deltas, rays, T_thresh):
        sigmas = sigmas.contiguous()
        rgbs = rgbs.contiguous()
        M, N, C = sigmas.shape[0], rays.shape[0], rgbs.shape[1]
        ws = torch.empty(N, device=sigmas.device)
        depth = torch.empty(N, device=sigmas.device)
        image = torch.empty(N, C, device=sigmas.device)
        ref().rm.composite_rays_train_forward(sigmas, rgbs, deltas, rays, M, N, C, T_thresh, False, ws, depth, image)
        ctx.save_for_backward(sigmas, rgbs, deltas, rays, ws, image)
        ctx.dims = (M, N, C, T_thresh)
        return ws, depth, image

    @staticmethod
    @custom_bwd(device_type='cuda')
    def backward(ctx, g_ws, g_depth, g_image):
        sigmas, rgbs, deltas, rays, ws, image = ctx.saved_tensors
        M, N, C, T_thresh = ctx.dims
        gs, gr, buf = torch.zeros_like(sigmas), torch.zeros_like(rgbs), torch.zeros_like(image)
        ref().rm.composite_rays_train_backward(g_ws.contiguous(), g_image.contiguous(), sigmas, rgbs, deltas, rays, False, ws, image,
                                               M, N, C, T_thresh, gs, gr, buf)
        return gs, gr, None, None, None


class RefGridEncode(Function):
    @staticmethod
    @custom_fwd(device_type='cuda')
    def forward(ctx, inputs, embeddings, offsets, S, H):
        inputs = inputs.contiguous()
        B, D = inputs.shape
        L = offsets.shape[0] - 1
        C = embeddings.shape[1]
        if torch.is_autocast_enabled('cuda'):
            embeddings = embeddings.to(torch.half)
        outputs = torch.empty(L, B, C, device=inputs.device, dtype=embeddings.dtype)
        dy_dx = torch.empty(1, device=inputs.device, dtype=embeddings.dtype)
        ref().ge.grid_encode_forward(inputs, embeddings, offsets, outputs, B, D, C, L, S, H, False, dy_dx, 0, True, 0)
        outputs = outputs.permute(1, 0, 2).reshape(B, L * C)
        ctx.save_for_backward(inputs, embeddings, offsets, dy_dx)
        ctx.dims = (B, D, C, L, S, H)
        return outputs

    @staticmethod
    @custom_bwd(device_type='cuda')
    def backward(ctx, grad):
        inputs, embeddings, offsets, dy_dx = ctx.saved_tensors
        B, D, C, L, S, H = ctx.dims
        grad = grad.view(B, L, C).permute(1, 0, 2).contiguous()
        ge = torch.zeros_like(embeddings)
        gi = torch.zeros(1, device=inputs.device, dtype=embeddings.dtype)
        ref().ge.grid_encode_backward(grad, inputs, embeddings, offsets, ge, B, D, C, L, S, H, False, dy_dx, gi, 0, True, 0)
        return None, ge, None, None, None


class RefGrid(nn.Module):
    def __init__(self, enc):
        super().__init__()
        self.embeddings = nn.Parameter(enc.embeddings.detach().clone())
        self.register_buffer('offsets', enc.offsets.clone())
        self.S = float(np.log2(enc.per_level_scale))
        self.H = enc.base_resolution

    def forward(self, x):
        x = (x + 1) / 2
        return RefGridEncode.apply(x, self.embeddings, self.offsets, self.S, self.H)


class TcnnStandIn(nn.Module):
    """torch fp16 linear chain with tcnn's padding; fp16 output [B, n_out]."""

    def __init__(self, net):
        super().__init__()
        self.n_in, self.n_out = net.n_input_dims, net.n_output_dims
        self.in_pad = net.in_pad
        self.params = nn.Parameter(net.params.detach().clone())
        self.shapes = net.layer_shapes
        self.sigmoid = net.out_act == 2

    def forward(self, x):
        B = x.shape[0]
        Bp = (B + 127) // 128 * 128
        h = x.to(torch.float)                                  # tcnn's binding casts the input to float
        h = F.pad(h, (0, self.in_pad - self.n_in, 0, Bp - B)).to(torch.half)
        p = self.params.to(torch.half)
        o = 0
        for i, (r, c) in enumerate(self.shapes):
            W = p[o:o + r * c].view(r, c)
            o += r * c
            h = F.linear(h, W)
            if i + 1 < len(self.shapes):
                h = F.relu(h)
        if self.sigmoid:
            h = torch.sigmoid(h)
        return h[:B, :self.n_out]


class _TruncExp(Function):
    @staticmethod
    @custom_fwd(device_type='cuda', cast_inputs=torch.float32)
    def forward(ctx, x):
        ctx.save_for_backward(x)
        return torch.exp(x)

    @staticmethod
    @custom_bwd(device_type='cuda')
    def backward(ctx, g):
        return g * torch.exp(ctx.saved_tensors[0].clamp(-15, 15))


class RefModel(nn.Module):
    def __init__(self, m):
        super().__init__()
        self.register_buffer('bbox_min', m.bbox_min.clone())
        self.register_buffer('bbox_size', m.bbox_size.clone())
        self.class_dim = m.class_dim
        self.x_density_embedder = RefGrid(m.x_density_embedder)
        self.x_color_embedder = RefGrid(m.x_color_embedder)
        self.density_net = TcnnStandIn(m.density_net)
        self.color1_net = TcnnStandIn(m.color1_net)
        self.color2_net = TcnnStandIn(m.color2_net)
        self.class_net = TcnnStandIn(m.class_net)

    def _forward(self, pts, dirs=None):
        pts = (pts - self.bbox_min) / self.bbox_size
        sigmas = _TruncExp.apply(self.density_net(self.x_density_embedder(pts)))
        if dirs is None:
            return sigmas
        xc = self.x_color_embedder(pts)
        classes = self.class_net(xc)
        rgbs = self.color2_net(self.color1_net(xc))
        return torch.cat((rgbs, classes), dim=1), sigmas

    def forward(self, pts, dirs=None, bsize=1000000):
        N = len(pts)
        if N < bsize:
            return self._forward(pts, dirs)
        sigmas = torch.empty((N, 1), device=pts.device)
        rgbs = torch.empty((N, 3 + self.class_dim), device=pts.device) if dirs is not None else None
        for s in range(0, N, bsize):
            e = min(N, s + bsize)
            if dirs is None:
                sigmas[s:e] = self._forward(pts[s:e])
            else:
                r, sg = self._forward(pts[s:e], dirs[s:e])
                rgbs[s:e] = r
                sigmas[s:e] = sg
        return sigmas if dirs is None else (rgbs, sigmas)


def time_reference_ext(n_rays, steps, device, warmup=3):
    """Train-step rays/s of the reference CUDA-extension path (same workload, same box)."""
    from bench import make_batches, unpack, N_CLASSES
    from nerfstyle_b200 import model as M, raymarching
    ref()
    torch.manual_seed(0)
    ours = M.StyleTCNerf([-2., -2., -2.], [2., 2., 2.], class_dim=N_CLASSES).to(device)
    model = RefModel(ours).to(device)
    # occupancy bitfield: one update_state of our renderer on the same random-init field (setup, untimed)
    r = M.Renderer(ours, 2.0, raymarch_channels=3 + N_CLASSES).to(device)
    with torch.autocast('cuda', dtype=torch.float16):
        r.update_state()
    bitfield = r.density_bitfield
    aabb = r.aabb
    del r, ours
    params = list(model.parameters())
    opt = torch.optim.Adam(params, lr=0.01, betas=(0.9, 0.999), eps=1e-15)
    scaler = torch.amp.GradScaler('cuda')
    ema = [p.detach().clone() for p in params]
    host, devb = make_batches(warmup + steps, n_rays, 0, 1, device)
    counter = torch.zeros(2, dtype=torch.int32, device=device)
    n_samples = 0

    def step(pack):
        nonlocal n_samples
        o, d, tgt, cls = unpack(pack)
        with torch.autocast('cuda', dtype=torch.float16):
            nears, fars = ref_near_far(o, d, aabb, 0.2)
            counter.zero_()
            xyzs, dirs, deltas, rays = ref_march_rays_train(o, d, 2.0, bitfield, 2, 128, nears, fars, counter)
            n_samples = xyzs.shape[0]
            rgbs, sigmas = model(xyzs, dirs)
            ws, depth, image = RefComposite.apply(sigmas, rgbs, deltas, rays, 1e-4)
            classes = image[:, 3:]
            img = image[:, :3] + (1 - ws).unsqueeze(-1)
            loss = torch.mean((img - tgt) ** 2) + 0.001 * F.cross_entropy(classes, cls)
        opt.zero_grad()
        scaler.scale(loss).backward()
        scaler.step(opt)
        scaler.update()
        with torch.no_grad():
            torch._foreach_mul_(ema, 0.95)
            torch._foreach_add_(ema, params, alpha=0.05)
        return loss

    for s in range(warmup):
        step(devb[s])
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for s in range(warmup, warmup + steps):
        loss = step(devb[s])
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / steps
    return {'value': round(n_rays / (ms / 1e3), 1), 'unit': 'rays/s', 'ms_per_step': round(ms, 3), 'steps': steps,
            'samples_per_step_last': int(n_samples), 'final_loss': float(loss),
            'what': "reference's raymarching + gridencoder CUDA extensions rebuilt for sm_100a (oracle/_ref) with the "
                    "reference's host-side allocation pattern, torch fp16 F.linear stand-in for tiny-cuda-nn, "
                    "torch Adam + EMA; device-resident inputs, no occupancy update inside the timed steps"}
