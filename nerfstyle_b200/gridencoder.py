"""Drop-in for the reference's `gridencoder` package (gridencoder/grid.py).

`GridEncoder` keeps the reference's constructor, attributes, parameter / buffer names (`embeddings`,
`offsets`), initialisation and forward signature; `grid_encode` is the autograd Function behind it.  The
CUDA side writes the point-major [B, L*C] layout directly (no permute copy) and takes point-major grads.
"""
import numpy as np
import torch
import torch.nn as nn
from torch.autograd import Function
from torch.amp import custom_bwd, custom_fwd

from . import _lib as L
from . import optim as _optim

_gridtype_to_id = {'hash': 0, 'tiled': 1}


def _require_complete_master(embeddings):
    """An fp32 read of a table whose optimizer shards it across ranks would see stale rows outside this rank's shard."""
    opt = _optim.optimizer_of(embeddings)
    if opt is not None and not opt.master_complete:
        raise RuntimeError('nerfstyle_b200.gridencoder: fp32 read of a hash table whose fp32 master copy is sharded across ranks '
                           '(only the fp16 copy is kept whole); call optimizer.gather_master() first or run under autocast')


class _grid_encode(Function):
    """grid.py:19-97 -> gridencoder.cu:439-494"""

    @staticmethod
    @custom_fwd(device_type='cuda')
    def forward(ctx, inputs, embeddings, offsets, per_level_scale, base_resolution, calc_grad_inputs=False,
                gridtype=0, align_corners=False, style=0):
        L.require_cuda(inputs, embeddings, offsets)
        if offsets.dtype != torch.int32:
            raise RuntimeError('offsets must be an int tensor')
        if not inputs.is_floating_point() or not embeddings.is_floating_point():
            raise RuntimeError('inputs / embeddings must be floating tensors')
        inputs = inputs.contiguous()
        if inputs.dtype != torch.float32:
            inputs = inputs.float()
        B, D = inputs.shape
        Lv = offsets.shape[0] - 1
        C = embeddings.shape[1]
        S = float(np.float32(np.log2(per_level_scale)))
        H = int(base_resolution)
        # grid.py:42-43: half-precision tables under autocast when C is even
        if torch.is_autocast_enabled('cuda') and C % 2 == 0:
            shadow = _optim.current_half_copy(embeddings)            # kept current by nerfstyle_b200.optim.FusedAdamEMA
            embeddings = shadow if shadow is not None else embeddings.to(torch.half)
        else:
            _require_complete_master(embeddings)
        embeddings = embeddings.contiguous()
        offsets = offsets.contiguous()
        dt = L.dtype_code(embeddings.dtype)
        outputs = torch.empty(B, Lv * C, device=inputs.device, dtype=embeddings.dtype)
        if calc_grad_inputs:
            dy_dx = torch.empty(B, Lv * D * C, device=inputs.device, dtype=embeddings.dtype)
        else:
            dy_dx = None
        L.Stats.units = B
        with torch.cuda.device(inputs.device):
            L.check(L.lib().nrf_grid_encode_forward(L.ptr(inputs), L.ptr(embeddings), L.ptr(offsets), L.ptr(outputs), B,
                                                    D, C, Lv, S, H, int(bool(calc_grad_inputs)), L.ptr(dy_dx),
                                                    int(gridtype), int(bool(align_corners)), int(style), dt, 1,
                                                    L.stream_of(inputs)), 'grid_encode_forward')
        ctx.save_for_backward(inputs, embeddings, offsets, dy_dx)
        ctx.dims = [B, D, C, Lv, S, H, gridtype]
        ctx.calc_grad_inputs = calc_grad_inputs
        ctx.align_corners = align_corners
        ctx.style = style
        return outputs

    @staticmethod
    @custom_bwd(device_type='cuda')
    def backward(ctx, grad):
        inputs, embeddings, offsets, dy_dx = ctx.saved_tensors
        B, D, C, Lv, S, H, gridtype = ctx.dims
        calc_grad_inputs = ctx.calc_grad_inputs
        grad = grad.contiguous()
        if grad.dtype != embeddings.dtype:
            grad = grad.to(embeddings.dtype)
        # fp16 grads (autocast) are accumulated into an fp32 table gradient: more accurate than the reference's
        # __half2 atomics (which swamp on the coarse levels) and it is the dtype the fp32 Parameter needs anyway
        grad_embeddings = torch.zeros(embeddings.shape, dtype=torch.float32, device=embeddings.device)
        grad_inputs = torch.zeros_like(inputs, dtype=embeddings.dtype) if calc_grad_inputs else None
        L.Stats.units = B
        with torch.cuda.device(inputs.device):
            L.check(L.lib().nrf_grid_encode_backward(L.ptr(grad), L.ptr(inputs), L.ptr(embeddings), L.ptr(offsets),
                                                     L.ptr(grad_embeddings), B, D, C, Lv, S, H,
                                                     int(bool(calc_grad_inputs)), L.ptr(dy_dx), L.ptr(grad_inputs),
                                                     int(gridtype), int(bool(ctx.align_corners)), int(ctx.style),
                                                     L.dtype_code(embeddings.dtype), L.DTYPE_F32, 1,
                                                     L.stream_of(inputs)),
                    'grid_encode_backward')
        if calc_grad_inputs:
            grad_inputs = grad_inputs.to(inputs.dtype)
            return grad_inputs, grad_embeddings, None, None, None, None, None, None, None
        return None, grad_embeddings, None, None, None, None, None, None, None


grid_encode = _grid_encode.apply


class _grid_encode_dual(Function):
    """Two encoders with identical geometry on the same points in one pass per direction (nrf_grid_encode_*_dual):
    cells, hash rows and trilinear weights are computed once.  Same values as two `_grid_encode` calls."""

    @staticmethod
    @custom_fwd(device_type='cuda')
    def forward(ctx, inputs, emb0, emb1, offsets, per_level_scale, base_resolution, gridtype, align_corners, style,
                xform=None, pipeline=False):
        L.require_cuda(inputs, emb0, emb1, offsets, xform)
        inputs = inputs.contiguous()
        if inputs.dtype != torch.float32:
            inputs = inputs.float()
        B, D = inputs.shape
        Lv = offsets.shape[0] - 1
        C = emb0.shape[1]
        if D != 3 or C != 2 or emb1.shape != emb0.shape:
            raise RuntimeError('grid_encode_dual: D=3, C=2 and equal table shapes are required')
        S = float(np.float32(np.log2(per_level_scale)))
        H = int(base_resolution)
        offsets = offsets.contiguous()
        # interleaved fp16 copies kept by the fused optimizer (optim.FusedAdamEMA, pair_tables): one gather serves both tables
        pair = None
        if torch.is_autocast_enabled('cuda'):
            pa, pb = _optim.current_half_pair(emb0), _optim.current_half_pair(emb1)
            if pa is not None and pb is not None and pa[0] is pb[0] and (pa[1], pb[1]) == (0, 1):
                pair = pa[0]
        ctx.sink = None
        sa, sb = _optim.live_grad_sink(emb0), _optim.live_grad_sink(emb1)
        if pair is not None and sa is not None and sb is not None and sa[0] is sb[0] and (sa[1], sb[1]) == (0, 1):
            ctx.sink = sa[0]
        L.Stats.units = B
        if pair is not None:
            out0 = torch.empty(B, Lv * C, device=inputs.device, dtype=pair.dtype)
            out1 = torch.empty_like(out0)
            esz = out0.element_size() * Lv * C
            chunks = _pipeline_chunks(B) if pipeline else [(0, B)]
            ready = []
            with torch.cuda.device(inputs.device):
                main = torch.cuda.current_stream(inputs.device)
                side = _side_stream(inputs.device) if len(chunks) > 1 else main
                if side is not main:
                    side.wait_stream(main)                   # inputs / tables / freshly allocated outputs are ordered on `main`
                with torch.cuda.stream(side):
                    for (r0, r1) in chunks:
                        L.check(L.lib().nrf_grid_encode_forward_pair(inputs.data_ptr() + 12 * r0, L.ptr(pair), L.ptr(offsets),
                                                                     out0.data_ptr() + esz * r0, out1.data_ptr() + esz * r0, r1 - r0, Lv, S, H,
                                                                     int(gridtype), int(bool(align_corners)), int(style),
                                                                     L.dtype_code(pair.dtype), L.ptr(xform), None, None, side.cuda_stream),
                                'grid_encode_forward_pair')
                        if side is not main:
                            ev = torch.cuda.Event()
                            ev.record(side)
                            ready.append((r0, r1, ev))
            # the consumer (tcnn.field_heads) makes `main` wait chunk by chunk, so the networks of chunk i run while chunk
            # i + 1 is still being gathered; grid_encode_dual() hands the events over (or waits for them itself)
            _last_ready[0] = ready
            ctx.save_for_backward(inputs, offsets, xform)
            ctx.meta = (B, Lv, S, H, gridtype, align_corners, style, pair.dtype, emb0.shape)
            return out0, out1
        embs = []
        for e in (emb0, emb1):
            if torch.is_autocast_enabled('cuda'):
                shadow = _optim.current_half_copy(e)
                e = shadow if shadow is not None else e.to(torch.half)
            else:
                _require_complete_master(e)
            embs.append(e.contiguous())
        if embs[0].dtype != embs[1].dtype:
            raise RuntimeError('grid_encode_dual: tables must share a dtype')
        dt = L.dtype_code(embs[0].dtype)
        out0 = torch.empty(B, Lv * C, device=inputs.device, dtype=embs[0].dtype)
        out1 = torch.empty_like(out0)
        with torch.cuda.device(inputs.device):
            L.check(L.lib().nrf_grid_encode_forward_dual(L.ptr(inputs), L.ptr(embs[0]), L.ptr(embs[1]), L.ptr(offsets),
                                                         L.ptr(out0), L.ptr(out1), B, Lv, S, H, int(gridtype),
                                                         int(bool(align_corners)), int(style), dt, L.ptr(xform),
                                                         L.stream_of(inputs)),
                    'grid_encode_forward_dual')
        ctx.save_for_backward(inputs, offsets, xform)
        ctx.meta = (B, Lv, S, H, gridtype, align_corners, style, embs[0].dtype, emb0.shape)
        return out0, out1

    @staticmethod
    @custom_bwd(device_type='cuda')
    def backward(ctx, g0, g1):
        inputs, offsets, xform = ctx.saved_tensors
        B, Lv, S, H, gridtype, align_corners, style, dtype, shape = ctx.meta
        dev = inputs.device
        need = [bool(ctx.needs_input_grad[1]), bool(ctx.needs_input_grad[2])]
        if not any(need):
            return None, None, None, None, None, None, None, None, None, None, None
        gs = []
        for g in (g0, g1):
            if g is None:
                g = torch.zeros(B, Lv * 2, dtype=dtype, device=dev)
            g = g.contiguous()
            gs.append(g if g.dtype == dtype else g.to(dtype))
        L.Stats.units = B
        if ctx.sink is not None and all(need):
            # the optimizer owns ONE interleaved f32 gradient buffer for both tables and reads it directly: the tables'
            # .grad stays None (FusedAdamEMA.grad_of); repeated backwards before a step accumulate in the buffer
            gp = ctx.sink.grad_pair_buffer()
            with torch.cuda.device(dev):
                L.check(L.lib().nrf_grid_encode_backward_pair(L.ptr(gs[0]), L.ptr(gs[1]), L.ptr(inputs), L.ptr(offsets), L.ptr(gp),
                                                              B, Lv, S, H, int(gridtype), int(bool(align_corners)), int(style),
                                                              L.dtype_code(dtype), L.ptr(xform), L.stream_of(inputs)),
                        'grid_encode_backward_pair')
            return None, None, None, None, None, None, None, None, None, None, None
        # a frozen table (requires_grad False, e.g. the density table of the stylization stage when the caller freezes it)
        # passes NULL: no zero fill, no reductions for it
        ge0 = torch.zeros(shape, dtype=torch.float32, device=dev) if need[0] else None
        ge1 = torch.zeros(shape, dtype=torch.float32, device=dev) if need[1] else None
        with torch.cuda.device(dev):
            L.check(L.lib().nrf_grid_encode_backward_dual(L.ptr(gs[0]) if need[0] else None, L.ptr(gs[1]) if need[1] else None,
                                                          L.ptr(inputs), L.ptr(offsets), L.ptr(ge0), L.ptr(ge1), B, Lv, S, H,
                                                          int(gridtype), int(bool(align_corners)), int(style), L.dtype_code(dtype),
                                                          L.DTYPE_F32, L.ptr(xform), L.stream_of(inputs)),
                    'grid_encode_backward_dual')
        return None, ge0, ge1, None, None, None, None, None, None, None, None


def same_geometry(a, b):
    """True when two GridEncoders address their tables identically (so one index computation serves both)."""
    return (a.input_dim == b.input_dim == 3 and a.level_dim == b.level_dim == 2 and a.num_levels == b.num_levels
            and a.per_level_scale == b.per_level_scale and a.base_resolution == b.base_resolution
            and a.gridtype_id == b.gridtype_id and a.align_corners == b.align_corners
            and a.embeddings.shape == b.embeddings.shape and torch.equal(a.offsets, b.offsets))


_last_ready = [None]   # [(row0, row1, event)] of the pipelined gather just launched by _grid_encode_dual.forward
_side = {}


def _side_stream(device):
    key = device.index if device.index is not None else torch.cuda.current_device()
    if key not in _side:
        _side[key] = torch.cuda.Stream(device=device)
    return _side[key]


# Chunks of the optional gather -> networks stream pipeline.  1 = off (default): measured on the B200, running the gather of
# chunk k + 1 next to the networks of chunk k does not pay -- both kernels are bound by the same SM resources (issue slots /
# L1TEX: 73 % + 26 % and 64 % + 35 % in ncu), so each slows the other down by what the overlap gains (train step 4.43 ms
# with 1 chunk, 4.48 with 4, 4.52 with 8, 4.67 with 16).  Kept for experiments (tools/pipe_sweep.py).
PIPELINE_CHUNKS = [1]


def _pipeline_chunks(B, min_rows=1 << 19, n=None):
    """Row ranges (multiples of 128 rows) for the gather -> networks pipeline; one range when the batch is small."""
    n = PIPELINE_CHUNKS[0] if n is None else n
    if B < 2 * min_rows:
        return [(0, B)]
    n = min(n, B // min_rows)
    per = (B // n + 127) // 128 * 128
    out, r = [], 0
    while r < B:
        out.append((r, min(B, r + per)))
        r += per
    return out


def take_ready(t):
    """The per-chunk completion events of an encoding produced with pipeline=True (and forget them), or None."""
    ev = getattr(t, '_nrf_ready', None)
    if ev is not None:
        del t._nrf_ready
    return ev


def wait_ready(*tensors):
    """Make the current stream wait for every gather chunk of these encodings (consumers that do not pipeline)."""
    for t in tensors:
        for (_, _, e) in (take_ready(t) or []):
            torch.cuda.current_stream(t.device).wait_event(e)


def grid_encode_dual(inputs, enc_a, enc_b, bound=1, style=0, xform=None, pipeline=False):
    """(enc_a(inputs, bound, style), enc_b(inputs, bound, style)) in one pass; the encoders must satisfy same_geometry().
    xform (device f32[7] = bbox_min, bbox_size, bound): `inputs` are RAW points and the kernel applies
    ((x - min) / size + bound) / (2 bound) itself (same f32 operations as the two Python lines it replaces)."""
    if xform is None:
        inputs = (inputs + bound) / (2 * bound)          # GridEncoder.forward, grid.py:174
    prefix_shape = list(inputs.shape[:-1])
    inputs = inputs.reshape(-1, enc_a.input_dim)
    _last_ready[0] = None
    o0, o1 = _grid_encode_dual.apply(inputs, enc_a.embeddings, enc_b.embeddings, enc_a.offsets, enc_a.per_level_scale,
                                     enc_a.base_resolution, enc_a.gridtype_id, enc_a.align_corners, style, xform, pipeline)
    ready, _last_ready[0] = _last_ready[0], None
    if ready:
        if pipeline and len(prefix_shape) == 1:
            o0._nrf_ready = o1._nrf_ready = ready        # consumed (and waited for) by tcnn.field_heads
            return o0, o1
        for (_, _, e) in ready:
            torch.cuda.current_stream(o0.device).wait_event(e)
    return o0.view(prefix_shape + [enc_a.output_dim]), o1.view(prefix_shape + [enc_b.output_dim])


class GridEncoder(nn.Module):
    """grid.py:103-191 (same arguments, attributes and state-dict keys)."""

    def __init__(self, input_dim=3, num_levels=16, level_dim=2, per_level_scale=2, base_resolution=16,
                 log2_hashmap_size=19, desired_resolution=None, gridtype='hash', align_corners=False):
        super().__init__()
        if desired_resolution is not None:
            per_level_scale = np.exp2(np.log2(desired_resolution / base_resolution) / (num_levels - 1))
        self.input_dim = input_dim
        self.num_levels = num_levels
        self.level_dim = level_dim
        self.per_level_scale = per_level_scale
        self.log2_hashmap_size = log2_hashmap_size
        self.base_resolution = base_resolution
        self.output_dim = num_levels * level_dim
        self.gridtype = gridtype
        self.gridtype_id = _gridtype_to_id[gridtype]
        self.align_corners = align_corners
        self.n_output_dims = num_levels * level_dim

        offsets = []
        offset = 0
        self.max_params = 2 ** log2_hashmap_size
        for i in range(num_levels):
            resolution = int(np.ceil(base_resolution * per_level_scale ** i))
            params_in_level = min(self.max_params, (resolution if align_corners else resolution + 1) ** input_dim)
            params_in_level = int(np.ceil(params_in_level / 8) * 8)
            offsets.append(offset)
            offset += params_in_level
        offsets.append(offset)
        offsets = torch.from_numpy(np.array(offsets, dtype=np.int32))
        self.register_buffer('offsets', offsets)
        self.n_params = offsets[-1] * level_dim
        self.embeddings = nn.Parameter(torch.empty(offset, level_dim))
        self.reset_parameters()

    def reset_parameters(self):
        std = 1e-4
        self.embeddings.data.uniform_(-std, std)

    def initialize(self, ref_embeddings, ref_offsets, num_styles=64):
        """grid.py:154-164 -> gridencoder.cu:551-571"""
        Lv = self.offsets.shape[0] - 1
        S = float(np.float32(np.log2(self.per_level_scale)))
        H = int(self.base_resolution)
        tmp = torch.zeros_like(self.embeddings)
        ref_embeddings = ref_embeddings.detach().float().contiguous()
        L.require_cuda(ref_embeddings, tmp, ref_offsets, self.offsets)
        with torch.cuda.device(tmp.device):
            L.check(L.lib().nrf_grid_initialize(L.ptr(ref_embeddings), L.ptr(tmp), L.ptr(ref_offsets.contiguous()),
                                                L.ptr(self.offsets), Lv, S, H, int(num_styles), L.stream_of(tmp)),
                    'grid_initialize')
        state_dict = self.state_dict()
        state_dict['embeddings'] = tmp
        self.load_state_dict(state_dict)

    def __repr__(self):
        return (f"GridEncoder: input_dim={self.input_dim} num_levels={self.num_levels} level_dim={self.level_dim} "
                f"resolution={self.base_resolution} -> "
                f"{int(round(self.base_resolution * self.per_level_scale ** (self.num_levels - 1)))} "
                f"per_level_scale={self.per_level_scale:.4f} params={tuple(self.embeddings.shape)} "
                f"gridtype={self.gridtype} align_corners={self.align_corners}")

    def forward(self, inputs, bound=1, style=0):
        inputs = (inputs + bound) / (2 * bound)  # map to [0, 1]
        prefix_shape = list(inputs.shape[:-1])
        inputs = inputs.view(-1, self.input_dim)
        outputs = grid_encode(inputs, self.embeddings, self.offsets, self.per_level_scale, self.base_resolution,
                              inputs.requires_grad, self.gridtype_id, self.align_corners, style)
        outputs = outputs.view(prefix_shape + [self.output_dim])
        return outputs
