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


class Timing:
    """CUDA-event brackets around every call into the reference's extensions (and the cuBLAS stand-in's forward)."""
    on = False
    events = []

    @classmethod
    def bracket(cls, name, fn):
        if not cls.on:
            return fn()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        out = fn()
        e1.record()
        cls.events.append((name, e0, e1))
        return out

    @classmethod
    def collect(cls, per):
        out = {}
        for name, a, b in cls.events:
            out[name] = out.get(name, 0.0) + a.elapsed_time(b) / per
        cls.events = []
        return {k: round(v, 4) for k, v in out.items()}


# host behaviour of the reference's wrappers: 'stock' = as shipped (zero-filled N*max_steps outputs + empty_cache() per call,
# raymarching.py:228-283); 'lean' = persistent output buffers, only the rows used by the previous call are re-zeroed, no
# empty_cache() -- isolates what the reference's KERNELS cost from what its host code costs
HOST_MODE = ['stock']
_lean = {}


def ref():
    if _RefOps.rm is None:
        _RefOps.rm = _load('_raymarching_ref')
        _RefOps.ge = _load('_gridencoder_ref')
    return _RefOps


def ref_near_far(o, d, aabb, min_near):
    N = o.shape[0]
    nears = torch.empty(N, device=o.device)
    fars = torch.empty(N, device=o.device)
    Timing.bracket('near_far_from_aabb', lambda: ref().rm.near_far_from_aabb(o, d, aabb, N, min_near, nears, fars))
    return nears, fars


def ref_march_rays_train(o, d, bound, bitfield, C, H, nears, fars, counter, align=128, max_steps=1024):
    N = o.shape[0]
    M = N * max_steps
    dev = o.device
    lean = HOST_MODE[0] == 'lean'
    if lean:
        st = _lean.get(('march', M))
        if st is None:
            st = _lean[('march', M)] = {'xyzs': torch.zeros(M, 3, device=dev), 'dirs': torch.zeros(M, 3, device=dev),
                                        'deltas': torch.zeros(M, 4, device=dev), 'used': 0, 'noises': torch.zeros(N, device=dev)}
        xyzs, dirs, deltas, noises = st['xyzs'], st['dirs'], st['deltas'], st['noises']
        if st['used']:
            xyzs[:st['used']].zero_(); dirs[:st['used']].zero_(); deltas[:st['used']].zero_()
    else:
        xyzs = torch.zeros(M, 3, device=dev)
        dirs = torch.zeros(M, 3, device=dev)
        deltas = torch.zeros(M, 4, device=dev)
        noises = torch.zeros(N, device=dev)
    rays = torch.empty(N, 3, dtype=torch.int32, device=dev)
    zh = torch.tensor((), device=dev)
    Timing.bracket('march_rays_train', lambda: ref().rm.march_rays_train(o, d, zh, bitfield, bound, 0.0, max_steps, False, N, C, H, M, nears,
                                                                         fars, xyzs, dirs, deltas, rays, counter, noises))
    m = counter[0].item()
    m += align - m % align
    if lean:
        st['used'] = m
    xyzs, dirs, deltas = xyzs[:m], dirs[:m], deltas[:m]
    if not lean:
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
        Timing.bracket('composite_rays_train_forward',
                       lambda: ref().rm.composite_rays_train_forward(sigmas, rgbs, deltas, rays, M, N, C, T_thresh, False, ws, depth, image))
        ctx.save_for_backward(sigmas, rgbs, deltas, rays, ws, image)
        ctx.dims = (M, N, C, T_thresh)
        return ws, depth, image

    @staticmethod
    @custom_bwd(device_type='cuda')
    def backward(ctx, g_ws, g_depth, g_image):
        sigmas, rgbs, deltas, rays, ws, image = ctx.saved_tensors
        M, N, C, T_thresh = ctx.dims
        gs, gr, buf = torch.zeros_like(sigmas), torch.zeros_like(rgbs), torch.zeros_like(image)
        gw, gi = g_ws.contiguous(), g_image.contiguous()
        Timing.bracket('composite_rays_train_backward',
                       lambda: ref().rm.composite_rays_train_backward(gw, gi, sigmas, rgbs, deltas, rays, False, ws, image, M, N, C, T_thresh,
                                                                      gs, gr, buf))
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
        Timing.bracket('grid_encode_forward',
                       lambda: ref().ge.grid_encode_forward(inputs, embeddings, offsets, outputs, B, D, C, L, S, H, False, dy_dx, 0, True, 0))
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
        Timing.bracket('grid_encode_backward',
                       lambda: ref().ge.grid_encode_backward(grad, inputs, embeddings, offsets, ge, B, D, C, L, S, H, False, dy_dx, gi, 0, True, 0))
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
        return Timing.bracket('mlp_cublas_forward', lambda: self._chain(x))

    def _chain(self, x):
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


def ref_update_state(model, density_grid, bitfield, local_step, H=128, cascade=2, bound=2.0, decay=0.95, thresh=10.0,
                     update_thres=256):
    """Renderer.update_state (renderer.py:139-194) on the reference's morton3D / morton3D_invert / packbits kernels."""
    dev = density_grid.device
    tmp = -torch.ones_like(density_grid)

    def sig(xyzs, cas):
        b = min(2 ** cas, bound)
        hg = b / H
        p = xyzs * (b - hg)
        p += (torch.rand_like(p) * 2 - 1) * hg
        return model(p).reshape(-1).detach()
    if local_step < update_thres:
        ax = torch.arange(H, dtype=torch.int32, device=dev)
        xx, yy, zz = torch.meshgrid(ax, ax, ax, indexing='ij')
        coords = torch.cat([xx.reshape(-1, 1), yy.reshape(-1, 1), zz.reshape(-1, 1)], dim=-1).contiguous()
        idx = torch.empty(coords.shape[0], dtype=torch.int32, device=dev)
        ref().rm.morton3D(coords, coords.shape[0], idx)
        xyzs = 2 * coords.float() / (H - 1) - 1
        for cas in range(cascade):
            tmp[cas, idx.long()] = sig(xyzs, cas)
    else:
        N = H ** 3 // 4
        for cas in range(cascade):
            coords = torch.randint(0, H, (N, 3), device=dev).int().contiguous()
            idx = torch.empty(N, dtype=torch.int32, device=dev)
            ref().rm.morton3D(coords, N, idx)
            occ = torch.nonzero(density_grid[cas] > 0).squeeze(-1)
            occ = occ[torch.randint(0, occ.shape[0], [N], dtype=torch.long, device=dev)].int().contiguous()
            occ_coords = torch.empty(N, 3, dtype=torch.int32, device=dev)
            ref().rm.morton3D_invert(occ, N, occ_coords)
            indices = torch.cat([idx, occ]).long()
            xyzs = 2 * torch.cat([coords, occ_coords]).float() / (H - 1) - 1
            tmp[cas, indices] = sig(xyzs, cas)
    valid = (density_grid >= 0) & (tmp >= 0)
    density_grid[valid] = torch.maximum(density_grid[valid] * decay, tmp[valid])
    mean = torch.mean(density_grid.clamp(min=0)).item()
    ref().rm.packbits(density_grid.contiguous(), bitfield.numel(), min(mean, thresh), bitfield)
    return mean


def time_reference_ext(n_rays, steps, device, warmup=3, update_iter=16):
    """Train-step rays/s of the reference CUDA-extension path (same workload, same box, occupancy update every 16 steps
    like bench.py's own arm).  Reported twice -- with the reference's stock host behaviour and with a lean host (see
    HOST_MODE) -- plus the per-kernel CUDA-event times of the reference's five kernels and of the cuBLAS stand-in."""
    from bench import make_batches, unpack, N_CLASSES
    from nerfstyle_b200 import model as M
    ref()
    host, devb = make_batches(warmup + steps, n_rays, 0, 1, device)
    out = {}
    for mode in ('stock', 'lean'):
        HOST_MODE[0] = mode
        _lean.clear()
        torch.manual_seed(0)
        ours = M.StyleTCNerf([-2., -2., -2.], [2., 2., 2.], class_dim=N_CLASSES).to(device)
        model = RefModel(ours).to(device)
        aabb = torch.tensor([-2., -2., -2., 2., 2., 2.], device=device)
        del ours
        grid = torch.zeros(2, 128 ** 3, device=device)
        bitfield = torch.zeros(2 * 128 ** 3 // 8, dtype=torch.uint8, device=device)
        params = list(model.parameters())
        opt = torch.optim.Adam(params, lr=0.01, betas=(0.9, 0.999), eps=1e-15)
        scaler = torch.amp.GradScaler('cuda')
        ema = [p.detach().clone() for p in params]
        counter = torch.zeros(2, dtype=torch.int32, device=device)
        state = {'n_samples': 0, 'it': 0}

        def step(pack):
            o, d, tgt, cls = unpack(pack)
            with torch.autocast('cuda', dtype=torch.float16):
                if state['it'] % update_iter == 0:
                    with torch.no_grad():
                        Timing.bracket('update_state', lambda: ref_update_state(model, grid, bitfield, state['it']))
                nears, fars = ref_near_far(o, d, aabb, 0.2)
                counter.zero_()
                xyzs, dirs, deltas, rays = ref_march_rays_train(o, d, 2.0, bitfield, 2, 128, nears, fars, counter)
                state['n_samples'] = xyzs.shape[0]
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
            state['it'] += 1
            return loss

        for s_ in range(warmup):
            step(devb[s_])
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for s_ in range(warmup, warmup + steps):
            loss = step(devb[s_])
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / steps
        out[mode] = {'value': round(n_rays / (ms / 1e3), 1), 'unit': 'rays/s', 'ms_per_step': round(ms, 3),
                     'samples_per_step_last': int(state['n_samples']), 'final_loss': float(loss.detach())}
        if mode == 'lean':        # per-kernel times in a separate instrumented pass (same state, 4 more steps)
            Timing.on, Timing.events = True, []
            for s_ in range(4):
                step(devb[s_])
            torch.cuda.synchronize()
            Timing.on = False
            out['kernels_ms_per_step'] = Timing.collect(4)
        del model, opt, ema, params
    HOST_MODE[0] = 'stock'
    _lean.clear()
    k = out['kernels_ms_per_step']
    res = dict(out['stock'])
    res.update({'steps': steps, 'kernels_only': out['lean'], 'kernels_ms_per_step': k,
                'reference_kernel_sum_ms': round(sum(v for n, v in k.items() if n not in ('mlp_cublas_forward', 'update_state')), 3),
                'what': "reference's raymarching + gridencoder CUDA extensions rebuilt for sm_100a (oracle/_ref), torch fp16 F.linear "
                        "stand-in for tiny-cuda-nn, torch Adam + EMA, occupancy update (renderer.py:139-194 on the reference's "
                        "morton3D / packbits kernels) every 16 steps; device-resident inputs.  `value` = the reference's stock host "
                        "behaviour (zero-filled N*max_steps outputs + torch.cuda.empty_cache() per march call); `kernels_only` = the "
                        "same kernels behind a lean host (persistent buffers, no empty_cache); `kernels_ms_per_step` = CUDA events "
                        "around each extension call (grid_encode_* are the two encoders together; the stand-in's backward runs "
                        "inside autograd and is not bracketed)"})
    return res


def time_reference_ext_render(w, h, device, frames=3, density_scale=50.0):
    """Full-frame inference (renderer.py:237-293 loop) on the reference's march_rays / composite_rays kernels + the cuBLAS
    stand-in, same field / occupancy / density_scale as bench.py's render_full_frame extra."""
    from bench import N_CLASSES
    from nerfstyle_b200 import model as M, raymarching, scenes
    ref()
    torch.manual_seed(0)
    ours = M.StyleTCNerf([-2., -2., -2.], [2., 2., 2.], class_dim=N_CLASSES).to(device)
    model = RefModel(ours).to(device)
    del ours
    bitfield = raymarching.packbits(scenes.analytic_density_grid(2, 128, 2.0).to(device), 0.5)
    aabb = torch.tensor([-2., -2., -2., 2., 2., 2.], device=device)
    intr = scenes.scaled_intrinsics(w, h)
    poses = scenes.synthetic_poses(frames + 1, 1)
    idx = torch.arange(0, w * h, device=device)
    Cch = 3 + N_CLASSES
    ms = []
    zh = torch.tensor((), device=device)
    for f in range(frames + 1):
        o, d = scenes.generate_rays(poses[f], intr, device, idx)
        torch.cuda.synchronize()
        import time
        t0 = time.perf_counter()
        with torch.no_grad(), torch.autocast('cuda', dtype=torch.float16):
            N = o.shape[0]
            nears, fars = ref_near_far(o, d, aabb, 0.2)
            ws = torch.zeros(N, device=device); depth = torch.zeros(N, device=device); image = torch.zeros(N, Cch, device=device)
            alive = torch.arange(N, dtype=torch.int32, device=device)
            rays_t = nears.clone()[:, None]
            step = 0
            while step < 1024:
                n_alive = len(alive)
                if n_alive <= 0:
                    break
                n_step = max(min(N // n_alive, 8), 1)
                Mp = n_alive * n_step
                Mp += 128 - (Mp % 128)
                xyzs = torch.zeros(Mp, 3, device=device); dirs = torch.zeros(Mp, 3, device=device); deltas = torch.zeros(Mp, 4, device=device)
                noises = torch.zeros(n_alive, device=device)
                ref().rm.march_rays(n_alive, n_step, alive, rays_t, o, d, zh, 2.0, 0.0, 1024, False, 2, 128, bitfield, nears, fars, xyzs, dirs,
                                    deltas, noises)
                rgbs, sigmas = model(xyzs, dirs)
                sigmas = (sigmas * density_scale).float().contiguous()
                ref().rm.composite_rays(n_alive, n_step, 1e-4, alive, rays_t, sigmas, rgbs.float().contiguous(), deltas, Cch, False, ws, depth, image)
                alive = alive[alive >= 0]
                step += n_step
            img = image[:, :3] + (1 - ws).unsqueeze(-1)
        float(img.sum().item())
        if f > 0:
            ms.append((time.perf_counter() - t0) * 1e3)
    t = sum(ms) / len(ms)
    return {'w': w, 'h': h, 'ms_per_frame': round(t, 2), 'mrays_per_s': round(w * h / t / 1e3, 2),
            'what': "renderer.py:237-293 loop on the reference's march_rays / composite_rays kernels (oracle/_ref) + cuBLAS stand-in"}
