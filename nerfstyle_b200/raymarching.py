"""Drop-in for the reference's `raymarching` package (raymarching/raymarching.py).

Same names, positional signatures, defaults, dtypes and in-place behaviour as the reference's
`Function.apply` objects, so `renderer.py` runs unchanged (`import raymarching`, see dropin.py).
Every op launches hand-written sm_100a kernels through the C ABI in include/nerfstyle_b200.h on the
current torch stream; there is no CPU path.
"""
import torch
from torch.autograd import Function
from torch.amp import custom_bwd, custom_fwd

from . import _lib as L

__all__ = ['near_far_from_aabb', 'sph_from_ray', 'morton3D', 'morton3D_invert', 'packbits', 'march_rays_train',
           'composite_rays_train', 'march_rays', 'composite_rays', 'march_rays_unbounded_train', 'compact_rays_alive', 'compact_rays_alive_nosync']


def _cuda(t):
    return t if t.is_cuda else t.cuda()


def _f32c(t, device=None):
    if device is not None and t.device != device:
        t = t.to(device)
    if t.dtype != torch.float32:
        t = t.float()
    return t.contiguous()


class _near_far_from_aabb(Function):
    """raymarching.py:19-52 -> kernel raymarching.cu:191-244"""

    @staticmethod
    @custom_fwd(device_type='cuda', cast_inputs=torch.float32)
    def forward(ctx, rays_o, rays_d, aabb, min_near=0.2):
        rays_o = _f32c(_cuda(rays_o)).view(-1, 3)
        rays_d = _f32c(_cuda(rays_d)).view(-1, 3)
        aabb = _f32c(aabb, rays_o.device)
        N = rays_o.shape[0]
        nears = torch.empty(N, dtype=rays_o.dtype, device=rays_o.device)
        fars = torch.empty(N, dtype=rays_o.dtype, device=rays_o.device)
        with torch.cuda.device(rays_o.device):
            L.check(L.lib().nrf_near_far_from_aabb(L.ptr(rays_o), L.ptr(rays_d), L.ptr(aabb), N, float(min_near),
                                                   L.ptr(nears), L.ptr(fars), L.stream_of(rays_o)),
                    'near_far_from_aabb')
        return nears, fars


near_far_from_aabb = _near_far_from_aabb.apply


class _sph_from_ray(Function):
    """raymarching.py:55-86 -> raymarching.cu:262-297"""

    @staticmethod
    @custom_fwd(device_type='cuda', cast_inputs=torch.float32)
    def forward(ctx, rays_o, rays_d, radius):
        rays_o = _f32c(_cuda(rays_o)).view(-1, 3)
        rays_d = _f32c(_cuda(rays_d)).view(-1, 3)
        N = rays_o.shape[0]
        coords = torch.empty(N, 2, dtype=rays_o.dtype, device=rays_o.device)
        with torch.cuda.device(rays_o.device):
            L.check(L.lib().nrf_sph_from_ray(L.ptr(rays_o), L.ptr(rays_d), float(radius), N, L.ptr(coords),
                                             L.stream_of(rays_o)), 'sph_from_ray')
        return coords


sph_from_ray = _sph_from_ray.apply


class _morton3D(Function):
    """raymarching.py:89-113 -> raymarching.cu:313-325"""

    @staticmethod
    def forward(ctx, coords):
        coords = _cuda(coords)
        N = coords.shape[0]
        indices = torch.empty(N, dtype=torch.int32, device=coords.device)
        coords = coords.int().contiguous()
        with torch.cuda.device(coords.device):
            L.check(L.lib().nrf_morton3D(L.ptr(coords), N, L.ptr(indices), L.stream_of(coords)), 'morton3D')
        return indices


morton3D = _morton3D.apply


class _morton3D_invert(Function):
    """raymarching.py:116-136 -> raymarching.cu:336-353"""

    @staticmethod
    def forward(ctx, indices):
        indices = _cuda(indices)
        N = indices.shape[0]
        coords = torch.empty(N, 3, dtype=torch.int32, device=indices.device)
        indices = indices.int().contiguous()
        with torch.cuda.device(indices.device):
            L.check(L.lib().nrf_morton3D_invert(L.ptr(indices), N, L.ptr(coords), L.stream_of(indices)),
                    'morton3D_invert')
        return coords


morton3D_invert = _morton3D_invert.apply


class _packbits(Function):
    """raymarching.py:139-167 -> raymarching.cu:367-388 (writes `bitfield` in place when given)"""

    @staticmethod
    @custom_fwd(device_type='cuda', cast_inputs=torch.float32)
    def forward(ctx, grid, thresh, bitfield=None):
        grid = _f32c(_cuda(grid))
        C = grid.shape[0]
        H3 = grid.shape[1]
        N = C * H3 // 8
        if bitfield is None:
            bitfield = torch.empty(N, dtype=torch.uint8, device=grid.device)
        L.require_cuda(bitfield)
        if bitfield.dtype != torch.uint8 or not bitfield.is_contiguous() or bitfield.numel() < N:
            raise RuntimeError('packbits: bitfield must be a contiguous uint8 tensor with C*H^3/8 entries')
        with torch.cuda.device(grid.device):
            L.check(L.lib().nrf_packbits(L.ptr(grid), N, float(thresh), L.ptr(bitfield), L.stream_of(grid)), 'packbits')
        return bitfield


packbits = _packbits.apply


STAGED_MARCH = True       # False: count pass + second DDA walk (nrf_march_rays_train_write), the round-1 form
_stage = {}


def _stage_buffer(dev, n):
    """Per-device [N * max_steps] f32 staging buffer of the emitted samples' marching times (33 MB at 8192 rays)."""
    key = dev.index if dev.index is not None else torch.cuda.current_device()
    buf = _stage.get(key)
    if buf is None or buf.numel() < n:
        buf = _stage[key] = torch.empty(n, dtype=torch.float32, device=dev)
    return buf


class _march_rays_train(Function):
    """raymarching.py:174-288 -> raymarching.cu:411-589.

    Differences that are invisible at this surface: offsets are the exclusive scan of the per-ray counts
    in ray order (deterministic; one valid schedule of the reference's racing atomicAdd), and when the
    true sample count is read back (force_all_rays or mean_count <= 0 -- the D2H sync is part of the
    reference contract, :275-281) the outputs are allocated at their final padded size instead of
    zero-filling N*max_steps rows and slicing."""

    @staticmethod
    @custom_fwd(device_type='cuda', cast_inputs=torch.float32)
    def forward(ctx, rays_o, rays_d, z_hats, bound, density_bitfield, C, H, nears, fars, step_counter=None,
                mean_count=-1, perturb=False, align=-1, force_all_rays=False, dt_gamma=0, max_steps=1024,
                is_ndc=False):
        rays_o = _f32c(_cuda(rays_o)).view(-1, 3)
        rays_d = _f32c(_cuda(rays_d)).view(-1, 3)
        dev = rays_o.device
        density_bitfield = _cuda(density_bitfield).contiguous()
        nears = _f32c(nears, dev)
        fars = _f32c(fars, dev)
        if is_ndc:
            z_hats = _f32c(_cuda(z_hats)).view(-1)
        else:
            z_hats = None
        N = rays_o.shape[0]
        M = N * max_steps
        if not force_all_rays and mean_count > 0:
            if align > 0:
                mean_count += align - mean_count % align
            M = mean_count
        rays = torch.empty(N, 3, dtype=torch.int32, device=dev)
        if step_counter is None:
            step_counter = torch.zeros(2, dtype=torch.int32, device=dev)
        L.require_cuda(step_counter)
        noises = None  # perturb is hard-disabled in the reference (raymarching.py:247)
        lib = L.lib()
        scratch = L.scratch(dev, lib.nrf_march_scratch_bytes(N))
        # staged marching: the count pass records every emitted sample's time, the second pass streams them back instead of
        # walking the occupancy grid again (csrc/raymarching.cu: k_march_emit); NDC keeps the two-walk form
        staged = (not is_ndc) and STAGED_MARCH and N * max_steps <= (1 << 29)
        t_stage = _stage_buffer(dev, N * max_steps) if staged else None
        with torch.cuda.device(dev):
            st = L.stream_of(rays_o)
            counter_before = step_counter[:1].clone()
            if staged:
                rc = lib.nrf_march_rays_train_count_staged(L.ptr(rays_o), L.ptr(rays_d), L.ptr(density_bitfield), float(bound),
                                                           float(dt_gamma), int(max_steps), N, int(C), int(H), L.ptr(nears),
                                                           L.ptr(fars), L.ptr(noises), L.ptr(rays), L.ptr(step_counter),
                                                           L.ptr(scratch), L.ptr(t_stage), st)
                if rc == -2:          # NRF_E_UNSUPPORTED (thread-per-ray mode selected): two-walk form
                    staged = False
                else:
                    L.check(rc, 'march_rays_train(count, staged)')
            if not staged:
                L.check(lib.nrf_march_rays_train_count(L.ptr(rays_o), L.ptr(rays_d), L.ptr(density_bitfield), float(bound),
                                                       float(dt_gamma), int(max_steps), N, int(C), int(H), L.ptr(nears),
                                                       L.ptr(fars), L.ptr(noises), L.ptr(rays), L.ptr(step_counter),
                                                       L.ptr(scratch), st), 'march_rays_train(count)')
            if force_all_rays or mean_count <= 0:
                # the contract's D2H read (one copy: [count before, count after])
                base, m = torch.cat([counter_before, step_counter[:1]]).tolist()
                zero_from = m
                if align > 0:
                    m += align - m % align
                rows = min(m, M) if M > 0 else 0
            else:
                rows, zero_from, base = M, 0, 0
            xyzs = torch.empty(rows, 3, dtype=torch.float32, device=dev)
            dirs = torch.empty(rows, 3, dtype=torch.float32, device=dev)
            deltas = torch.empty(rows, 4, dtype=torch.float32, device=dev)
            if rows > 0 and zero_from == 0:
                xyzs.zero_(), dirs.zero_(), deltas.zero_()
                zero_from = rows
            elif rows > 0 and base > 0:   # caller did not zero the counter: rows below the base stay zero
                xyzs[:base].zero_(), dirs[:base].zero_(), deltas[:base].zero_()
            if N > 0 and rows > 0 and staged:
                L.check(lib.nrf_march_rays_train_emit(L.ptr(rays_o), L.ptr(rays_d), float(bound), float(dt_gamma), int(max_steps), N,
                                                      int(C), int(H), M, rows, min(zero_from, rows), L.ptr(nears), L.ptr(noises),
                                                      L.ptr(rays), L.ptr(t_stage), L.ptr(xyzs), L.ptr(dirs), L.ptr(deltas), st),
                        'march_rays_train(emit)')
            elif N > 0 and rows > 0:
                L.check(lib.nrf_march_rays_train_write(L.ptr(rays_o), L.ptr(rays_d), L.ptr(z_hats),
                                                       L.ptr(density_bitfield), float(bound), float(dt_gamma),
                                                       int(max_steps), int(bool(is_ndc)), N, int(C), int(H), M, rows,
                                                       min(zero_from, rows), L.ptr(nears), L.ptr(fars), L.ptr(noises),
                                                       L.ptr(rays), L.ptr(xyzs), L.ptr(dirs), L.ptr(deltas), st),
                        'march_rays_train(write)')
            if (force_all_rays or mean_count <= 0) and base == 0 and rows >= zero_from:
                # every row below `zero_from` belongs to exactly one ray (offsets are the exclusive scan of the counts) and the
                # rows above are padding: composite_rays_train's backward can then skip the reference's zero fill
                rays._nrf_dense_rows = int(zero_from)
        return xyzs, dirs, deltas, rays


march_rays_train = _march_rays_train.apply


def march_rays_unbounded_train(*args, **kwargs):
    """Exported by the reference's extension (bindings.cpp:20) but its kernel returns before writing any
    sample (raymarching.cu:707) and no Python caller exists; kept as a name only."""
    raise NotImplementedError('march_rays_unbounded_train is dead code in the reference (raymarching.cu:707)')


class _composite_rays_train(Function):
    """raymarching.py:291-350 -> raymarching.cu:807-879 (fwd), :905-986 (bwd)"""

    @staticmethod
    @custom_fwd(device_type='cuda', cast_inputs=torch.float32)
    def forward(ctx, sigmas, rgbs, deltas, rays, T_thresh=1e-4, is_ndc=False):
        # the kernels are fp32: cast explicitly (custom_fwd only casts while autocast is enabled)
        sigmas = _f32c(sigmas)
        rgbs = _f32c(rgbs)
        deltas = _f32c(deltas)
        rays = rays.contiguous()
        L.require_cuda(sigmas, rgbs, deltas, rays)
        M = sigmas.shape[0]
        N = rays.shape[0]
        C = rgbs.shape[1]
        dev = sigmas.device
        weights_sum = torch.empty(N, dtype=sigmas.dtype, device=dev)
        depth = torch.empty(N, dtype=sigmas.dtype, device=dev)
        image = torch.empty(N, C, dtype=sigmas.dtype, device=dev)
        with torch.cuda.device(dev):
            L.check(L.lib().nrf_composite_rays_train_forward(L.ptr(sigmas), L.ptr(rgbs), L.ptr(deltas), L.ptr(rays), M,
                                                             N, C, float(T_thresh), int(bool(is_ndc)),
                                                             L.ptr(weights_sum), L.ptr(depth), L.ptr(image),
                                                             L.stream_of(sigmas)), 'composite_rays_train_forward')
        ctx.save_for_backward(sigmas, rgbs, deltas, rays, weights_sum, depth, image)
        dense = getattr(rays, '_nrf_dense_rows', None)
        ctx.dense_rows = dense if (dense is not None and C <= 32 and dense <= M) else None
        ctx.dims = [M, N, C, T_thresh]
        ctx.is_ndc = is_ndc
        return weights_sum, depth, image

    @staticmethod
    @custom_bwd(device_type='cuda')
    def backward(ctx, grad_weights_sum, grad_depth, grad_image):
        # grad_depth is not propagated (raymarching.py:331)
        grad_weights_sum = _f32c(grad_weights_sum)
        grad_image = _f32c(grad_image)
        sigmas, rgbs, deltas, rays, weights_sum, depth, image = ctx.saved_tensors
        M, N, C, T_thresh = ctx.dims
        if ctx.dense_rows is not None:
            # rays from this package's march_rays_train partition rows [0, dense_rows) exactly: the kernel writes every one of
            # them (gradient or zero) and only the <= 128 padding rows are filled here, instead of the reference's two full
            # zero fills (raymarching.py:339-340; 187 MB per step at 3.9 M samples)
            grad_sigmas = torch.empty_like(sigmas)
            grad_rgbs = torch.empty_like(rgbs)
            if ctx.dense_rows < M:
                grad_sigmas[ctx.dense_rows:].zero_()
                grad_rgbs[ctx.dense_rows:].zero_()
            write_zeros = 1
        else:
            grad_sigmas = torch.zeros_like(sigmas)
            grad_rgbs = torch.zeros_like(rgbs)
            write_zeros = 0
        with torch.cuda.device(sigmas.device):
            L.check(L.lib().nrf_composite_rays_train_backward_ex(L.ptr(grad_weights_sum), L.ptr(grad_image), L.ptr(sigmas),
                                                                 L.ptr(rgbs), L.ptr(deltas), L.ptr(rays),
                                                                 int(bool(ctx.is_ndc)), L.ptr(weights_sum), L.ptr(image),
                                                                 M, N, C, float(T_thresh), L.ptr(grad_sigmas),
                                                                 L.ptr(grad_rgbs), write_zeros, L.stream_of(sigmas)),
                    'composite_rays_train_backward')
        return grad_sigmas, grad_rgbs, None, None, None, None


composite_rays_train = _composite_rays_train.apply


class _march_rays(Function):
    """raymarching.py:357-427 -> raymarching.cu:1005-1120"""

    @staticmethod
    @custom_fwd(device_type='cuda', cast_inputs=torch.float32)
    def forward(ctx, n_alive, n_step, rays_alive, rays_t, rays_o, rays_d, z_hats, bound, density_bitfield, C, H, near,
                far, align=-1, perturb=False, dt_gamma=0, max_steps=1024, is_ndc=False):
        rays_o = _f32c(_cuda(rays_o)).view(-1, 3)
        rays_d = _f32c(_cuda(rays_d)).view(-1, 3)
        dev = rays_o.device
        if is_ndc:
            z_hats = _f32c(_cuda(z_hats)).view(-1)
        else:
            z_hats = None
        L.require_cuda(rays_alive, rays_t, density_bitfield, near, far)
        n_alive, n_step = int(n_alive), int(n_step)
        M = n_alive * n_step
        if align > 0:
            M += align - (M % align)
        xyzs = torch.empty(M, 3, dtype=torch.float32, device=dev)
        dirs = torch.empty(M, 3, dtype=torch.float32, device=dev)
        deltas = torch.empty(M, 4, dtype=torch.float32, device=dev)
        noises = torch.rand(n_alive, dtype=torch.float32, device=dev) if perturb else None
        if M > 0:
            with torch.cuda.device(dev):
                L.check(L.lib().nrf_march_rays(n_alive, n_step, L.ptr(rays_alive), L.ptr(rays_t), L.ptr(rays_o),
                                               L.ptr(rays_d), L.ptr(z_hats), float(bound), float(dt_gamma),
                                               int(max_steps), int(bool(is_ndc)), int(C), int(H),
                                               L.ptr(density_bitfield), L.ptr(near), L.ptr(far), L.ptr(xyzs),
                                               L.ptr(dirs), L.ptr(deltas), L.ptr(noises), M, 1, L.stream_of(rays_o)),
                        'march_rays')
        return xyzs, dirs, deltas


march_rays = _march_rays.apply


class _composite_rays(Function):
    """raymarching.py:430-462 -> raymarching.cu:1134-1231 (in place; returns an empty tuple)"""

    @staticmethod
    @custom_fwd(device_type='cuda', cast_inputs=torch.float32)
    def forward(ctx, n_alive, n_step, rays_alive, rays_t, sigmas, rgbs, deltas, is_ndc, weights_sum, depth, image,
                T_thresh=1e-2):
        t_size = 2 if is_ndc else 1
        assert rays_t.shape[-1] == t_size
        C = rgbs.shape[-1]
        sigmas = _f32c(sigmas)
        rgbs = _f32c(rgbs)
        deltas = _f32c(deltas)
        L.require_cuda(rays_alive, rays_t, sigmas, rgbs, deltas, weights_sum, depth, image)
        for name, t in (('rays_t', rays_t), ('weights_sum', weights_sum), ('depth', depth), ('image', image)):
            if t.dtype != torch.float32:
                raise RuntimeError('composite_rays: %s must be float32' % name)
        if rays_alive.dtype != torch.int32:
            raise RuntimeError('composite_rays: rays_alive must be int32')
        for name, t in (('rays_alive', rays_alive), ('rays_t', rays_t), ('weights_sum', weights_sum), ('depth', depth),
                        ('image', image)):
            if not t.is_contiguous():
                raise RuntimeError('composite_rays: %s is updated in place and must be contiguous' % name)
        with torch.cuda.device(sigmas.device):
            L.check(L.lib().nrf_composite_rays(int(n_alive), int(n_step), float(T_thresh), L.ptr(rays_alive),
                                               L.ptr(rays_t), L.ptr(sigmas), L.ptr(rgbs), L.ptr(deltas), C,
                                               int(bool(is_ndc)), L.ptr(weights_sum), L.ptr(depth), L.ptr(image),
                                               L.stream_of(sigmas)), 'composite_rays')
        return tuple()


composite_rays = _composite_rays.apply


def compact_rays_alive(rays_alive):
    """Extension: `rays_alive[rays_alive >= 0]` (renderer.py:284) as one stable compaction kernel.
    Returns (compacted tensor, n_alive); the count is read back (the caller's loop needs it on the host)."""
    rays_alive = rays_alive.contiguous()
    L.require_cuda(rays_alive)
    n = rays_alive.shape[0]
    out = torch.empty_like(rays_alive)
    cnt = torch.empty(1, dtype=torch.int32, device=rays_alive.device)
    lib = L.lib()
    scratch = L.scratch(rays_alive.device, lib.nrf_march_scratch_bytes(n))
    with torch.cuda.device(rays_alive.device):
        L.check(lib.nrf_compact_alive(L.ptr(rays_alive), n, L.ptr(out), L.ptr(cnt), L.ptr(scratch),
                                      L.stream_of(rays_alive)), 'compact_alive')
    k = int(cnt.item())
    return out[:k], k


def compact_rays_alive_nosync(rays_alive, out=None, cnt=None):
    """Same compaction WITHOUT the read-back: returns (out, cnt) where out has the input's length, its first cnt[0]
    entries are the surviving ray ids in order and the rest are -1 (dead slots, skipped by march_rays /
    composite_rays); cnt is a device int32[1].  Lets an inference loop refresh its host-side ray count only every few
    iterations."""
    rays_alive = rays_alive.contiguous()
    L.require_cuda(rays_alive)
    n = rays_alive.shape[0]
    if out is None:
        out = torch.empty_like(rays_alive)
    if cnt is None:
        cnt = torch.empty(1, dtype=torch.int32, device=rays_alive.device)
    lib = L.lib()
    scratch = L.scratch(rays_alive.device, lib.nrf_march_scratch_bytes(n))
    with torch.cuda.device(rays_alive.device):
        L.check(lib.nrf_compact_alive(L.ptr(rays_alive), n, L.ptr(out), L.ptr(cnt), L.ptr(scratch),
                                      L.stream_of(rays_alive)), 'compact_alive')
    return out, cnt
