"""numpy-facing wrappers over oracle.c (TEST INFRASTRUCTURE ONLY -- see oracle/__init__.py).

The wrappers restate the *Python* half of the reference operators as well (allocation, padding and
slicing rules of raymarching/raymarching.py and gridencoder/grid.py), so that tests can compare the
CUDA drop-in modules against them call for call.
"""
import ctypes
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SRC = os.path.join(_HERE, 'oracle.c')
_SO = os.path.join(_HERE, 'liboracle.so')
_lib = None

c_f = ctypes.POINTER(ctypes.c_float)
c_i = ctypes.POINTER(ctypes.c_int32)
c_u8 = ctypes.POINTER(ctypes.c_uint8)
c_u32 = ctypes.POINTER(ctypes.c_uint32)
u32 = ctypes.c_uint32
f32 = ctypes.c_float
i32 = ctypes.c_int


def build(force=False):
    """Compile oracle.c -> liboracle.so with gcc (no FP contraction; explicit fmaf only)."""
    if not force and os.path.exists(_SO) and os.path.getmtime(_SO) >= os.path.getmtime(_SRC):
        return _SO
    cmd = ['gcc', '-O2', '-std=gnu11', '-ffp-contract=off', '-fno-fast-math', '-fopenmp', '-fPIC', '-shared',
           '-fvisibility=hidden', _SRC, '-o', _SO, '-lm']
    subprocess.check_call(cmd)
    return _SO


def lib():
    global _lib
    if _lib is None:
        build()
        _lib = ctypes.CDLL(_SO)
        _lib.orc_fast_hash3.restype = u32
        _lib.orc_fast_hash3.argtypes = [u32] * 4
        _lib.orc_grid_index3.restype = u32
        _lib.orc_grid_index3.argtypes = [u32] * 8
        _lib.orc_level_resolution.restype = u32
        _lib.orc_level_resolution.argtypes = [u32, f32, u32]
    return _lib


def _f(a):
    return a.ctypes.data_as(c_f)


def _i(a):
    return a.ctypes.data_as(c_i)


def _b(a):
    return a.ctypes.data_as(c_u8)


def _v(a):
    return a.ctypes.data_as(ctypes.c_void_p)


def _np(x, dtype=np.float32):
    if hasattr(x, 'detach'):
        x = x.detach().cpu().numpy()
    return np.ascontiguousarray(x, dtype=dtype)


# ----------------------------------------------------------------------------- raymarching utils
def fast_hash3(x, y, z, style=0):
    return int(lib().orc_fast_hash3(x, y, z, style))


def grid_index3(C, gridtype, hashmap_size, resolution, x, y, z, style=0):
    return int(lib().orc_grid_index3(C, gridtype, hashmap_size, resolution, x, y, z, style))


def level_resolution(level, S, H):
    return int(lib().orc_level_resolution(level, np.float32(S), H))


def morton3D(coords):
    coords = _np(coords, np.int32).reshape(-1, 3)
    out = np.empty(coords.shape[0], np.int32)
    lib().orc_morton3D(_i(coords), u32(coords.shape[0]), _i(out))
    return out


def morton3D_invert(indices):
    indices = _np(indices, np.int32).reshape(-1)
    out = np.empty((indices.shape[0], 3), np.int32)
    lib().orc_morton3D_invert(_i(indices), u32(indices.shape[0]), _i(out))
    return out


def packbits(grid, thresh, bitfield=None):
    """raymarching.py:139-167"""
    grid = _np(grid)
    C, H3 = grid.shape
    N = C * H3 // 8
    if bitfield is None:
        bitfield = np.empty(N, np.uint8)
    lib().orc_packbits(_f(grid), u32(N), f32(thresh), _b(bitfield))
    return bitfield


def near_far_from_aabb(rays_o, rays_d, aabb, min_near=0.2):
    """raymarching.py:19-52"""
    rays_o = _np(rays_o).reshape(-1, 3)
    rays_d = _np(rays_d).reshape(-1, 3)
    aabb = _np(aabb)
    N = rays_o.shape[0]
    nears = np.empty(N, np.float32)
    fars = np.empty(N, np.float32)
    lib().orc_near_far_from_aabb(_f(rays_o), _f(rays_d), _f(aabb), u32(N), f32(min_near), _f(nears), _f(fars))
    return nears, fars


def sph_from_ray(rays_o, rays_d, radius):
    rays_o = _np(rays_o).reshape(-1, 3)
    rays_d = _np(rays_d).reshape(-1, 3)
    N = rays_o.shape[0]
    coords = np.empty((N, 2), np.float32)
    lib().orc_sph_from_ray(_f(rays_o), _f(rays_d), f32(radius), u32(N), _f(coords))
    return coords


# ----------------------------------------------------------------------------- training ops
def march_rays_train_count(rays_o, rays_d, bound, bitfield, C, H, nears, fars, dt_gamma=0., max_steps=1024,
                           noises=None):
    rays_o = _np(rays_o).reshape(-1, 3)
    rays_d = _np(rays_d).reshape(-1, 3)
    bitfield = _np(bitfield, np.uint8)
    nears, fars = _np(nears), _np(fars)
    N = rays_o.shape[0]
    counts = np.empty(N, np.int32)
    nz = _f(_np(noises)) if noises is not None else None
    lib().orc_march_rays_train_count(_f(rays_o), _f(rays_d), _b(bitfield), f32(bound), f32(dt_gamma), u32(max_steps),
                                     u32(N), u32(C), u32(H), _f(nears), _f(fars), nz, _i(counts))
    return counts


def march_rays_train(rays_o, rays_d, z_hats, bound, density_bitfield, C, H, nears, fars, step_counter=None,
                     mean_count=-1, perturb=False, align=-1, force_all_rays=False, dt_gamma=0, max_steps=1024,
                     is_ndc=False, noises=None):
    """raymarching.py:174-288 (perturb is hard-disabled there, :247; `noises` lets tests inject some)."""
    rays_o = _np(rays_o).reshape(-1, 3)
    rays_d = _np(rays_d).reshape(-1, 3)
    bitfield = _np(density_bitfield, np.uint8)
    nears, fars = _np(nears), _np(fars)
    N = rays_o.shape[0]
    M = N * max_steps
    if not force_all_rays and mean_count > 0:
        if align > 0:
            mean_count += align - mean_count % align
        M = mean_count
    xyzs = np.zeros((M, 3), np.float32)
    dirs = np.zeros((M, 3), np.float32)
    deltas = np.zeros((M, 4), np.float32)
    rays = np.empty((N, 3), np.int32)
    if step_counter is None:
        step_counter = np.zeros(2, np.int32)
    nz = np.zeros(N, np.float32) if noises is None else _np(noises)
    zh = _np(z_hats).reshape(-1) if is_ndc else np.zeros(1, np.float32)
    lib().orc_march_rays_train(_f(rays_o), _f(rays_d), _f(zh), _b(bitfield), f32(bound), f32(dt_gamma),
                               u32(max_steps), i32(int(is_ndc)), u32(N), u32(C), u32(H), u32(M), _f(nears), _f(fars),
                               _f(xyzs), _f(dirs), _f(deltas), _i(rays), _i(step_counter), _f(nz))
    if force_all_rays or mean_count <= 0:
        m = int(step_counter[0])
        if align > 0:
            m += align - m % align
        xyzs, dirs, deltas = xyzs[:m], dirs[:m], deltas[:m]
    return xyzs, dirs, deltas, rays


def composite_rays_train_forward(sigmas, rgbs, deltas, rays, T_thresh=1e-4, is_ndc=False):
    """raymarching.py:291-325"""
    sigmas = _np(sigmas).reshape(-1)
    rgbs = _np(rgbs)
    deltas = _np(deltas)
    rays = _np(rays, np.int32)
    M, N, C = sigmas.shape[0], rays.shape[0], rgbs.shape[1]
    ws = np.empty(N, np.float32)
    depth = np.empty(N, np.float32)
    image = np.empty((N, C), np.float32)
    lib().orc_composite_rays_train_forward(_f(sigmas), _f(rgbs), _f(deltas), _i(rays), u32(M), u32(N), u32(C),
                                           f32(T_thresh), i32(int(is_ndc)), _f(ws), _f(depth), _f(image))
    return ws, depth, image


def composite_rays_train_backward(grad_ws, grad_image, sigmas, rgbs, deltas, rays, weights_sum, image,
                                  T_thresh=1e-4, is_ndc=False):
    """raymarching.py:327-347"""
    sigmas = _np(sigmas).reshape(-1)
    rgbs, deltas, rays = _np(rgbs), _np(deltas), _np(rays, np.int32)
    grad_ws, grad_image, weights_sum, image = _np(grad_ws), _np(grad_image), _np(weights_sum), _np(image)
    M, N, C = sigmas.shape[0], rays.shape[0], rgbs.shape[1]
    assert C <= 64
    gs = np.zeros_like(sigmas)
    gr = np.zeros_like(rgbs)
    lib().orc_composite_rays_train_backward(_f(grad_ws), _f(grad_image), _f(sigmas), _f(rgbs), _f(deltas), _i(rays),
                                            i32(int(is_ndc)), _f(weights_sum), _f(image), u32(M), u32(N), u32(C),
                                            f32(T_thresh), _f(gs), _f(gr))
    return gs, gr


# ----------------------------------------------------------------------------- inference ops
def march_rays(n_alive, n_step, rays_alive, rays_t, rays_o, rays_d, z_hats, bound, density_bitfield, C, H, near, far,
               align=-1, perturb=False, dt_gamma=0, max_steps=1024, is_ndc=False, noises=None):
    """raymarching.py:357-427"""
    rays_o = _np(rays_o).reshape(-1, 3)
    rays_d = _np(rays_d).reshape(-1, 3)
    rays_alive = _np(rays_alive, np.int32)
    rays_t = _np(rays_t)
    bitfield = _np(density_bitfield, np.uint8)
    near, far = _np(near), _np(far)
    M = n_alive * n_step
    if align > 0:
        M += align - (M % align)
    xyzs = np.zeros((M, 3), np.float32)
    dirs = np.zeros((M, 3), np.float32)
    deltas = np.zeros((M, 4), np.float32)
    nz = np.zeros(max(n_alive, 1), np.float32) if noises is None else _np(noises)
    zh = _np(z_hats).reshape(-1) if is_ndc else np.zeros(1, np.float32)
    lib().orc_march_rays(u32(n_alive), u32(n_step), _i(rays_alive), _f(rays_t), _f(rays_o), _f(rays_d), _f(zh),
                         f32(bound), f32(dt_gamma), u32(max_steps), i32(int(is_ndc)), u32(C), u32(H), _b(bitfield),
                         _f(near), _f(far), _f(xyzs), _f(dirs), _f(deltas), _f(nz))
    return xyzs, dirs, deltas


def composite_rays(n_alive, n_step, rays_alive, rays_t, sigmas, rgbs, deltas, is_ndc, weights_sum, depth, image,
                   T_thresh=1e-2):
    """raymarching.py:430-462 -- in place on rays_alive, rays_t, weights_sum, depth, image (numpy arrays)."""
    sigmas = _np(sigmas).reshape(-1)
    rgbs, deltas = _np(rgbs), _np(deltas)
    for a, dt in ((rays_alive, np.int32), (rays_t, np.float32), (weights_sum, np.float32), (depth, np.float32),
                  (image, np.float32)):
        assert isinstance(a, np.ndarray) and a.dtype == dt and a.flags['C_CONTIGUOUS']
    C = rgbs.shape[-1]
    lib().orc_composite_rays(u32(n_alive), u32(n_step), f32(T_thresh), _i(rays_alive), _f(rays_t), _f(sigmas),
                             _f(rgbs), _f(deltas), u32(C), i32(int(is_ndc)), _f(weights_sum), _f(depth), _f(image))
    return tuple()


# ----------------------------------------------------------------------------- hash-grid encoder
def grid_offsets(input_dim=3, num_levels=16, level_dim=2, per_level_scale=2, base_resolution=16,
                 log2_hashmap_size=19, desired_resolution=None, align_corners=False):
    """gridencoder/grid.py:110-141 -- level offsets (rows) and the possibly overridden per_level_scale."""
    if desired_resolution is not None:
        per_level_scale = np.exp2(np.log2(desired_resolution / base_resolution) / (num_levels - 1))
    offsets, offset = [], 0
    max_params = 2 ** log2_hashmap_size
    for i in range(num_levels):
        resolution = int(np.ceil(base_resolution * per_level_scale ** i))
        params_in_level = min(max_params, (resolution if align_corners else resolution + 1) ** input_dim)
        params_in_level = int(np.ceil(params_in_level / 8) * 8)
        offsets.append(offset)
        offset += params_in_level
    offsets.append(offset)
    return np.array(offsets, dtype=np.int32), per_level_scale


def grid_encode_forward(inputs, embeddings, offsets, per_level_scale, base_resolution, calc_grad_inputs=False,
                        gridtype=0, align_corners=False, style=0, half=False, return_indices=False):
    """grid.py:22-66.  Returns outputs [B, L*C] (after the reference's permute), dy_dx or None[, idx [L,B,8]]."""
    inputs = _np(inputs)
    B, D = inputs.shape
    offsets = _np(offsets, np.int32)
    L = offsets.shape[0] - 1
    dt = np.float16 if half else np.float32
    embeddings = _np(embeddings, dt)
    C = embeddings.shape[1]
    S = np.float32(np.log2(per_level_scale))
    outputs = np.empty((L, B, C), dt)
    dy_dx = np.empty((B, L * D * C), dt) if calc_grad_inputs else np.empty(1, dt)
    idx = np.empty((L, B, 1 << D), np.uint32) if return_indices else None
    lib().orc_grid_encode_forward(_f(inputs), _v(embeddings), _i(offsets), _v(outputs), u32(B), u32(D), u32(C), u32(L),
                                  f32(S), u32(base_resolution), i32(int(calc_grad_inputs)), _v(dy_dx), u32(gridtype),
                                  i32(int(align_corners)), u32(style), i32(int(half)),
                                  idx.ctypes.data_as(c_u32) if idx is not None else None)
    out = np.ascontiguousarray(outputs.transpose(1, 0, 2).reshape(B, L * C))
    res = (out, dy_dx if calc_grad_inputs else None)
    if return_indices:
        res = res + (idx,)
    return res


def grid_encode_backward(grad, inputs, offsets, n_rows, C, per_level_scale, base_resolution, gridtype=0,
                         align_corners=False, style=0, half=False):
    """grid.py:71-97 -- grad [B, L*C] -> grad_embeddings [rows, C]."""
    inputs = _np(inputs)
    B, D = inputs.shape
    offsets = _np(offsets, np.int32)
    L = offsets.shape[0] - 1
    dt = np.float16 if half else np.float32
    grad = _np(grad, dt).reshape(B, L, C).transpose(1, 0, 2)
    grad = np.ascontiguousarray(grad)
    S = np.float32(np.log2(per_level_scale))
    ge = np.zeros((n_rows, C), dt)
    lib().orc_grid_encode_backward(_v(grad), _f(inputs), _i(offsets), _v(ge), u32(B), u32(D), u32(C), u32(L), f32(S),
                                   u32(base_resolution), u32(gridtype), i32(int(align_corners)), u32(style),
                                   i32(int(half)))
    return ge


def grid_input_backward(grad, dy_dx, B, D, C, L):
    grad = np.ascontiguousarray(_np(grad).reshape(B, L, C).transpose(1, 0, 2))
    dy_dx = _np(dy_dx)
    out = np.empty((B, D), np.float32)
    lib().orc_grid_input_backward(_f(grad), _f(dy_dx), _f(out), u32(B), u32(D), u32(C), u32(L))
    return out


def grid_initialize(ref_embeddings, ref_offsets, offsets, n_rows, per_level_scale, base_resolution, num_styles):
    ref_embeddings = _np(ref_embeddings)
    ref_offsets, offsets = _np(ref_offsets, np.int32), _np(offsets, np.int32)
    L = offsets.shape[0] - 1
    out = np.zeros((n_rows, 2), np.float32)
    lib().orc_grid_initialize(_f(ref_embeddings), _f(out), _i(ref_offsets), _i(offsets), u32(L),
                              f32(np.float32(np.log2(per_level_scale))), u32(base_resolution), u32(num_styles))
    return out
