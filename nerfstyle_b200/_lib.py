"""ctypes binding of libnerfstyle_b200.so (the C ABI declared in include/nerfstyle_b200.h).

There is no CPU fallback: if the library is missing or a call fails, the ops raise RuntimeError.
"""
import ctypes
import os

import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, 'libnerfstyle_b200.so')

_vp = ctypes.c_void_p
_u32 = ctypes.c_uint32
_i32 = ctypes.c_int
_f32 = ctypes.c_float
_u64 = ctypes.c_uint64

# name -> (restype, argtypes); mirrors include/nerfstyle_b200.h one to one
SIGNATURES = {
    'nrf_error_string': (ctypes.c_char_p, [_i32]),
    'nrf_last_cuda_error': (_i32, []),
    'nrf_version': (_i32, []),
    'nrf_device_info': (_i32, [_vp, _vp, _vp, _vp]),
    'nrf_near_far_from_aabb': (_i32, [_vp, _vp, _vp, _u32, _f32, _vp, _vp, _vp]),
    'nrf_sph_from_ray': (_i32, [_vp, _vp, _f32, _u32, _vp, _vp]),
    'nrf_morton3D': (_i32, [_vp, _u32, _vp, _vp]),
    'nrf_morton3D_invert': (_i32, [_vp, _u32, _vp, _vp]),
    'nrf_packbits': (_i32, [_vp, _u32, _f32, _vp, _vp]),
    'nrf_march_scratch_bytes': (_u64, [_u32]),
    'nrf_march_rays_train_count': (_i32, [_vp, _vp, _vp, _f32, _f32, _u32, _u32, _u32, _u32, _vp, _vp, _vp, _vp, _vp,
                                          _vp, _vp]),
    'nrf_march_rays_train_write': (_i32, [_vp, _vp, _vp, _vp, _f32, _f32, _u32, _i32, _u32, _u32, _u32, _u32, _u32,
                                          _u32, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp]),
    'nrf_march_rays_train_count_staged': (_i32, [_vp, _vp, _vp, _f32, _f32, _u32, _u32, _u32, _u32, _vp, _vp, _vp, _vp, _vp,
                                                 _vp, _vp, _vp]),
    'nrf_march_rays_train_emit': (_i32, [_vp, _vp, _f32, _f32, _u32, _u32, _u32, _u32, _u32, _u32, _u32, _vp, _vp, _vp, _vp, _vp, _vp,
                                         _vp, _vp]),
    'nrf_march_rays_train': (_i32, [_vp, _vp, _vp, _vp, _f32, _f32, _u32, _i32, _u32, _u32, _u32, _u32, _vp, _vp, _vp,
                                    _vp, _vp, _vp, _vp, _vp, _vp, _vp]),
    'nrf_composite_rays_train_forward': (_i32, [_vp, _vp, _vp, _vp, _u32, _u32, _u32, _f32, _i32, _vp, _vp, _vp, _vp]),
    'nrf_composite_rays_train_backward': (_i32, [_vp, _vp, _vp, _vp, _vp, _vp, _i32, _vp, _vp, _u32, _u32, _u32, _f32,
                                                 _vp, _vp, _vp]),
    'nrf_composite_rays_train_backward_ex': (_i32, [_vp, _vp, _vp, _vp, _vp, _vp, _i32, _vp, _vp, _u32, _u32, _u32, _f32,
                                                    _vp, _vp, _i32, _vp]),
    'nrf_march_rays': (_i32, [_u32, _u32, _vp, _vp, _vp, _vp, _vp, _f32, _f32, _u32, _i32, _u32, _u32, _vp, _vp, _vp,
                              _vp, _vp, _vp, _vp, _u32, _i32, _vp]),
    'nrf_composite_rays': (_i32, [_u32, _u32, _f32, _vp, _vp, _vp, _vp, _vp, _u32, _i32, _vp, _vp, _vp, _vp]),
    'nrf_compact_alive': (_i32, [_vp, _u32, _vp, _vp, _vp, _vp]),
    'nrf_grid_encode_forward': (_i32, [_vp, _vp, _vp, _vp, _u32, _u32, _u32, _u32, _f32, _u32, _i32, _vp, _u32, _i32,
                                       _u32, _i32, _i32, _vp]),
    'nrf_grid_encode_backward': (_i32, [_vp, _vp, _vp, _vp, _vp, _u32, _u32, _u32, _u32, _f32, _u32, _i32, _vp, _vp,
                                        _u32, _i32, _u32, _i32, _i32, _i32, _vp]),
    'nrf_grid_encode_forward_dual': (_i32, [_vp, _vp, _vp, _vp, _vp, _vp, _u32, _u32, _f32, _u32, _u32, _i32, _u32, _i32, _vp,
                                            _vp]),
    'nrf_grid_encode_backward_dual': (_i32, [_vp, _vp, _vp, _vp, _vp, _vp, _u32, _u32, _f32, _u32, _u32, _i32, _u32, _i32, _i32,
                                             _vp, _vp]),
    'nrf_grid_encode_forward_pair': (_i32, [_vp, _vp, _vp, _vp, _vp, _u32, _u32, _f32, _u32, _u32, _i32, _u32, _i32, _vp, _vp,
                                            _vp, _vp]),
    'nrf_grid_encode_backward_pair': (_i32, [_vp, _vp, _vp, _vp, _vp, _u32, _u32, _f32, _u32, _u32, _i32, _u32, _i32, _vp, _vp]),
    'nrf_grid_initialize': (_i32, [_vp, _vp, _vp, _vp, _u32, _f32, _u32, _u32, _vp]),
    'nrf_mlp_forward': (_i32, [_vp, _i32, _vp, _u32, _u32, _u32, _u32, _u32, _i32, _i32, _vp, _i32, _vp]),
    'nrf_mlp_backward': (_i32, [_vp, _i32, _vp, _vp, _i32, _u32, _u32, _u32, _u32, _u32, _i32, _i32, _f32, _vp, _i32,
                                _vp, _vp]),
    'nrf_mlp_forward_ex': (_i32, [_vp, _i32, _vp, _u32, _u32, _u32, _u32, _u32, _i32, _i32, _vp, _i32, _u32, _vp]),
    'nrf_mlp_backward_ex': (_i32, [_vp, _i32, _vp, _vp, _i32, _u32, _u32, _u32, _u32, _u32, _u32, _i32, _i32, _f32, _vp, _i32,
                                   _i32, _vp, _vp]),
    'nrf_field_forward': (_i32, [_vp, _vp, _vp, _vp, _vp, _vp, _u32, _u32, _vp, _vp, _u32, _vp, _vp, _vp]),
    'nrf_mlp_forward_f32': (_i32, [_vp, _i32, _vp, _u32, _u32, _u32, _u32, _u32, _i32, _i32, _vp, _u32, _vp]),
    'nrf_mlp_backward_f32': (_i32, [_vp, _i32, _vp, _vp, _u32, _u32, _u32, _u32, _u32, _u32, _i32, _i32, _vp, _i32, _vp, _vp]),
    'nrf_march_rays_dev': (_i32, [_vp, _u32, _vp, _vp, _vp, _vp, _f32, _f32, _u32, _u32, _u32, _vp, _vp, _vp, _vp, _vp, _vp]),
    'nrf_composite_rays_dev': (_i32, [_vp, _u32, _f32, _vp, _vp, _vp, _vp, _vp, _u32, _vp, _vp, _vp, _vp]),
    'nrf_compact_alive_dev': (_i32, [_vp, _u32, _vp, _vp, _vp, _vp]),
    'nrf_grid_encode_forward_dual_dev': (_i32, [_vp, _vp, _vp, _vp, _vp, _vp, _u32, _u32, _f32, _u32, _u32, _i32, _u32, _i32,
                                                _vp, _vp, _vp, _vp]),
    'nrf_mlp_forward_dev': (_i32, [_vp, _i32, _vp, _u32, _u32, _u32, _u32, _u32, _i32, _i32, _vp, _i32, _u32, _vp, _vp]),
    'nrf_sh_encode_forward': (_i32, [_vp, _u32, _u32, _vp, _i32, _vp]),
    'nrf_nnfm_forward': (_i32, [_vp, _vp, _u32, _u32, _u32, _vp, _vp, _vp, _u32, _vp, _vp, _vp, _vp]),
    'nrf_nnfm_scratch_bytes': (_u64, [_u32, _u32, _u32]),
    'nrf_opt_state_bytes': (_u64, []),
    'nrf_grads_check': (_i32, [_vp, _u64, _vp, _vp]),
    'nrf_adam_step': (_i32, [_vp, _vp, _vp, _vp, _vp, _vp, _u64, _vp, _f32, _f32, _f32, _f32, _f32, _f32, _vp]),
    'nrf_adam_step_ex': (_i32, [_vp, _vp, _vp, _vp, _vp, _vp, _u64, _vp, _f32, _f32, _f32, _f32, _f32, _f32, _u32, _u32, _vp]),
    'nrf_adam_step_pair': (_i32, [_vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _u64, _vp, _f32, _f32, _f32, _f32, _f32, _f32,
                                  _vp]),
    'nrf_scaler_update': (_i32, [_vp, _f32, _f32, _i32, _vp]),
    'nrf_small_allreduce_p2p': (_i32, [_vp, _u32, _u32, _vp, _vp, _vp]),
    'nrf_adam_step_pair_p2p': (_i32, [_vp, _vp, _vp, _vp, _vp, _vp, _u32, _u64, _vp, _vp, _vp, _vp, _vp, _vp, _u64, _vp, _f32, _f32,
                                      _f32, _f32, _f32, _f32, _vp]),
    'nrf_occ_points_full': (_i32, [_vp, _vp, _u32, _u32, _vp, _vp]),
    'nrf_occ_points_sparse': (_i32, [_vp, _vp, _vp, _vp, _vp, _vp, _vp, _u32, _u32, _u32, _vp, _vp]),
    'nrf_occ_flags': (_i32, [_vp, _u32, _u32, _vp, _vp]),
    'nrf_occ_scatter_max': (_i32, [_vp, _vp, _vp, _f32, _u32, _u32, _u32, _vp]),
    'nrf_occ_scratch_bytes': (_u64, []),
    'nrf_occ_update': (_i32, [_vp, _vp, _f32, _f32, _u64, _f32, _vp, _vp, _vp]),
    'nrf_packbits_dev': (_i32, [_vp, _u32, _vp, _vp, _vp]),
    'nrf_recon_loss_scratch_bytes': (_u64, [_u32]),
    'nrf_recon_loss': (_i32, [_vp, _vp, _vp, _vp, _u32, _u32, _f32, _vp, _vp, _vp, _vp, _vp]),
    'nrf_generate_rays': (_i32, [_vp, _f32, _f32, _f32, _f32, _u32, _u32, _u32, _u32, _vp, _i32, _vp, _u32, _u32, _vp, _vp,
                                 _vp, _vp]),
}
# tuning / extension entry points that are not part of the reference-replacing ABI
EXTRA_SIGNATURES = {
    'nrf_grid_set_tuning': (None, [_i32, _i32, _i32]),
    'nrf_grid_set_transpose_min': (None, [_i32]),
    'nrf_grid_set_bwd_walk': (None, [_i32]),
    'nrf_grid_set_bwd_walk_queue': (None, [_i32]),
    'nrf_march_set_mode': (None, [_i32]),
    'nrf_mlp_set_mode': (None, [_i32]),
    'nrf_mlp_get_mode': (_i32, []),
    'nrf_nnfm_set_mode': (None, [_i32]),
    'nrf_mlp_set_tuning': (None, [_i32, _i32]),
    'nrf_mlp_set_profile': (None, [_vp]),
}

DTYPE_F32, DTYPE_F16 = 0, 1
ACT = {'none': 0, 'relu': 1, 'sigmoid': 2, 'exponential': 3, 'trunc_exp': 4}

_lib = None

# kernels launched per C call (lower bound; used for the `gpu_launches` claim of bench.py)
KERNELS_PER_CALL = {
    'nrf_near_far_from_aabb': 1, 'nrf_sph_from_ray': 1, 'nrf_morton3D': 1, 'nrf_morton3D_invert': 1, 'nrf_packbits': 1,
    'nrf_march_rays_train_count': 4, 'nrf_march_rays_train_write': 1, 'nrf_march_rays_train_count_staged': 4, 'nrf_march_rays_train_emit': 1, 'nrf_march_rays_train': 4,
    'nrf_composite_rays_train_forward': 1, 'nrf_composite_rays_train_backward': 1, 'nrf_composite_rays_train_backward_ex': 1, 'nrf_march_rays': 1,
    'nrf_composite_rays': 1, 'nrf_compact_alive': 3, 'nrf_grid_encode_forward': 1, 'nrf_grid_encode_backward': 1, 'nrf_grid_encode_forward_dual': 1,
    'nrf_grid_encode_backward_dual': 1, 'nrf_grid_encode_forward_pair': 1, 'nrf_grid_encode_backward_pair': 1,
    'nrf_grid_initialize': 1, 'nrf_mlp_forward': 1, 'nrf_mlp_backward': 1, 'nrf_sh_encode_forward': 1, 'nrf_march_rays_dev': 1,
    'nrf_composite_rays_dev': 1, 'nrf_compact_alive_dev': 4, 'nrf_grid_encode_forward_dual_dev': 1, 'nrf_mlp_forward_dev': 1, 'nrf_mlp_forward_ex': 1, 'nrf_mlp_backward_ex': 1, 'nrf_mlp_forward_f32': 1, 'nrf_field_forward': 1, 'nrf_mlp_backward_f32': 1, 'nrf_nnfm_forward': 5,
    'nrf_adam_step': 1, 'nrf_adam_step_ex': 1, 'nrf_adam_step_pair': 1, 'nrf_adam_step_pair_p2p': 1, 'nrf_small_allreduce_p2p': 1, 'nrf_grads_check': 1, 'nrf_scaler_update': 1, 'nrf_generate_rays': 1,
    'nrf_occ_points_full': 1, 'nrf_occ_points_sparse': 1, 'nrf_occ_flags': 1, 'nrf_occ_scatter_max': 1, 'nrf_occ_update': 2,
    'nrf_packbits_dev': 1, 'nrf_recon_loss': 1,
}


class Stats:
    """Launch counter + optional CUDA-event timing of selected entry points (bench.py / profiling only)."""
    launches = 0
    calls = {}
    timed = set()          # entry-point names to bracket with CUDA events on the current stream
    events = []            # (name, start_event, end_event, units)
    units = 0              # set by the wrapper just before a timed call (e.g. number of points)

    @classmethod
    def reset(cls):
        cls.launches = 0
        cls.calls = {}
        cls.events = []


class _Proxy:
    def __init__(self, cdll):
        self._cdll = cdll
        self._cache = {}

    def __getattr__(self, name):
        fn = self._cache.get(name)
        if fn is None:
            raw = getattr(self._cdll, name)
            k = KERNELS_PER_CALL.get(name)
            if k is None:
                fn = raw
            else:
                def fn(*args, _raw=raw, _k=k, _name=name):
                    Stats.launches += _k
                    Stats.calls[_name] = Stats.calls.get(_name, 0) + 1
                    if _name in Stats.timed:
                        e0 = torch.cuda.Event(enable_timing=True)
                        e1 = torch.cuda.Event(enable_timing=True)
                        e0.record()
                        rc = _raw(*args)
                        e1.record()
                        Stats.events.append((_name, e0, e1, Stats.units))
                        return rc
                    return _raw(*args)
            self._cache[name] = fn
        return fn


def lib():
    """Load the CUDA library (once).  Raises RuntimeError when it has not been built."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise RuntimeError(
            'nerfstyle_b200: %s is missing -- build it with `python -m nerfstyle_b200.build` '
            '(there is no CPU fallback)' % LIB_PATH)
    l = ctypes.CDLL(LIB_PATH)
    for table in (SIGNATURES, EXTRA_SIGNATURES):
        for name, (res, args) in table.items():
            fn = getattr(l, name, None)
            if fn is None:
                raise RuntimeError('nerfstyle_b200: %s does not export %s (stale build?)' % (LIB_PATH, name))
            fn.restype = res
            fn.argtypes = args
    _lib = _Proxy(l)
    return _lib


def check(code, what):
    if code != 0:
        l = lib()
        msg = l.nrf_error_string(code).decode()
        if code == -3:
            msg += ' (cudaError %d)' % l.nrf_last_cuda_error()
        raise RuntimeError('nerfstyle_b200.%s failed: %s' % (what, msg))


def ptr(t):
    """Device pointer of a tensor (None -> NULL)."""
    return None if t is None else t.data_ptr()


def stream_of(t):
    return torch.cuda.current_stream(t.device).cuda_stream


def require_cuda(*tensors):
    for t in tensors:
        if t is not None and not t.is_cuda:
            raise RuntimeError('nerfstyle_b200: expected a CUDA tensor, got device %s' % t.device)


def dtype_code(dt):
    if dt == torch.float32:
        return DTYPE_F32
    if dt == torch.float16:
        return DTYPE_F16
    raise RuntimeError('nerfstyle_b200: unsupported dtype %s (float32 or float16)' % dt)


_scratch = {}


def scratch(device, nbytes):
    """Small per-device scratch buffer reused across calls on the same stream order."""
    key = (device.index if device.index is not None else torch.cuda.current_device())
    buf = _scratch.get(key)
    if buf is None or buf.numel() < nbytes:
        buf = torch.empty(max(int(nbytes), 1 << 16), dtype=torch.uint8, device=device)
        _scratch[key] = buf
    return buf
