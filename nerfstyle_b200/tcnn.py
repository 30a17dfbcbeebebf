"""tiny-cuda-nn style `Network` / `Encoding` modules on the hand-written fused MLP kernels.

Mirrors the part of the `tinycudann` torch binding the reference uses
(networks/style_nerf.py:34-98, networks/tcnn_nerf.py:38-50,87-122):

    tcnn.Network(n_input_dims=, n_output_dims=, network_config={'otype': 'FullyFusedMLP', 'activation': 'ReLU',
                 'output_activation': 'None'|'Sigmoid', 'n_neurons': 64, 'n_hidden_layers': h}, seed=)

* one flat fp32 `params` Parameter in tcnn's FullyFusedMLP layout (row-major [out,in] per layer, input and
  output widths padded to 16), cast to fp16 on every forward;
* fp16 output `[batch, n_output_dims]`; attributes `n_input_dims`, `n_output_dims`, `params`, `dtype`,
  `loss_scale`, `seed`.

tiny-cuda-nn is an un-vendored, unpinned dependency of the reference (README.md:25-26), so its numerics
are restated, not copied: fp16 operands, fp32 tensor-core accumulation, hidden activations rounded to
fp16 between layers (tcnn itself accumulates in fp16); weight init is Xavier-uniform from a torch
generator seeded with `seed` (tcnn's pcg32 stream is not reproducible here; parity tests inject weights).
"""
import math

import torch
import torch.nn as nn
from torch.autograd import Function

from . import _lib as L

_ACT = {'none': 0, 'relu': 1, 'sigmoid': 2, 'exponential': 3}


def _pad16(n):
    return (n + 15) // 16 * 16


class _mlp_function(Function):
    @staticmethod
    def forward(ctx, x, params, cfg, owner=None):
        n_in, n_out, n_hidden, width, hidden_act, out_act, loss_scale = cfg
        L.require_cuda(x, params)
        if x.dtype not in (torch.float16, torch.float32):
            x = x.float()
        x = x.contiguous()
        params_h = half_params(params, owner)
        B = x.shape[0]
        if params_h.dtype == torch.float32:                  # parity mode
            y = torch.empty(B, n_out, dtype=torch.float32, device=x.device)
            with torch.cuda.device(x.device):
                L.check(L.lib().nrf_mlp_forward_f32(L.ptr(x), L.dtype_code(x.dtype), L.ptr(params_h), B, n_in, n_out, n_hidden,
                                                    width, hidden_act, out_act, L.ptr(y), n_out, L.stream_of(x)),
                        'mlp_forward_f32')
            ctx.save_for_backward(x, params_h)
            ctx.cfg = cfg
            ctx.need_dx = ctx.needs_input_grad[0]
            ctx.need_dp = ctx.needs_input_grad[1]
            return y
        y = torch.empty(B, n_out, dtype=torch.float16, device=x.device)
        with torch.cuda.device(x.device):
            L.check(L.lib().nrf_mlp_forward(L.ptr(x), L.dtype_code(x.dtype), L.ptr(params_h), B, n_in, n_out, n_hidden,
                                            width, hidden_act, out_act, L.ptr(y), L.DTYPE_F16, L.stream_of(x)),
                    'mlp_forward')
        ctx.save_for_backward(x, params_h)
        ctx.cfg = cfg
        ctx.need_dx = ctx.needs_input_grad[0]
        ctx.need_dp = ctx.needs_input_grad[1]
        return y

    @staticmethod
    def backward(ctx, dy):
        x, params_h = ctx.saved_tensors
        n_in, n_out, n_hidden, width, hidden_act, out_act, loss_scale = ctx.cfg
        if dy.dtype not in (torch.float16, torch.float32):
            dy = dy.float()
        dy = dy.contiguous()
        B = x.shape[0]
        dx = torch.empty_like(x) if ctx.need_dx else None
        dparams = torch.zeros(params_h.shape, dtype=torch.float32, device=x.device) if ctx.need_dp else None
        if (ctx.need_dx or ctx.need_dp) and params_h.dtype == torch.float32:          # parity mode
            dy = dy.float().contiguous()
            with torch.cuda.device(x.device):
                L.check(L.lib().nrf_mlp_backward_f32(L.ptr(x), L.dtype_code(x.dtype), L.ptr(params_h), L.ptr(dy), n_out, B, n_in,
                                                     n_out, n_hidden, width, hidden_act, out_act, L.ptr(dx), 0, L.ptr(dparams),
                                                     L.stream_of(x)), 'mlp_backward_f32')
        elif ctx.need_dx or ctx.need_dp:
            with torch.cuda.device(x.device):
                L.check(L.lib().nrf_mlp_backward(L.ptr(x), L.dtype_code(x.dtype), L.ptr(params_h), L.ptr(dy),
                                                 L.dtype_code(dy.dtype), B, n_in, n_out, n_hidden, width, hidden_act,
                                                 out_act, float(loss_scale), L.ptr(dx), L.dtype_code(x.dtype),
                                                 L.ptr(dparams), L.stream_of(x)), 'mlp_backward')
        return dx, dparams, None, None


_parity = False


def set_parity_mode(on):
    """fp32 PARITY MODE: every Network evaluates with fp32 weights / activations / accumulation on the SIMT
    kernels of csrc/mlp_f32.cu and returns fp32; nothing is rounded to fp16.  For parity checks, not for speed."""
    global _parity
    _parity = bool(on)


class parity_mode:
    """`with tcnn.parity_mode(): ...` -- scoped set_parity_mode(True)."""

    def __enter__(self):
        self._prev = _parity
        set_parity_mode(True)
        return self

    def __exit__(self, *exc):
        set_parity_mode(self._prev)
        return False


_cache_epoch = 0          # > 0 and odd while a cache_half_params() scope is open


class cache_half_params:
    """Scope in which the fp16 casts of the networks' parameters are cached (an inference loop calls each network
    hundreds of times per frame with unchanged weights).  Outside such a scope every forward re-casts, exactly like
    tinycudann: a version-counter check alone would miss writes through `param.data` (torch_ema's copy_to / restore)."""

    def __enter__(self):
        global _cache_epoch
        self._outer = _cache_epoch
        if _cache_epoch % 2 == 0:
            _cache_epoch += 1
        return self

    def __exit__(self, *exc):
        global _cache_epoch
        if self._outer % 2 == 0:
            _cache_epoch += 1          # leaving the outermost scope invalidates everything cached inside it
        return False


def half_params(params, owner=None):
    """fp16 copy of a flat parameter vector for the kernels.  FusedAdamEMA keeps one current on the parameter
    (`_nrf_half_copy`, written by the optimizer kernel); inside a cache_half_params() scope the cast is cached on
    `owner`; otherwise it is made afresh."""
    if _parity:
        return params.detach().float().contiguous()          # parity mode: the kernels read the fp32 master weights
    from .optim import current_half_copy
    h = current_half_copy(params)          # re-cast automatically when the parameter was written since (checkpoint load)
    if h is not None:
        return h
    if owner is None or _cache_epoch % 2 == 0:
        return params.detach().to(torch.float16).contiguous()
    key = (_cache_epoch, params._version, params.data_ptr())
    if getattr(owner, '_half_key', None) != key:
        owner._half = params.detach().to(torch.float16).contiguous()
        owner._half_key = key
    return owner._half


def _fwd_ex(net, x, params_h, y, col, n_out, out_act):
    """y[:, col:col+n_out] = net(x) through the extended C entry point (y may be wider than n_out; f16 or f32)."""
    if params_h.dtype == torch.float32:                      # parity mode (y must be f32)
        L.check(L.lib().nrf_mlp_forward_f32(L.ptr(x), L.dtype_code(x.dtype), L.ptr(params_h), x.shape[0], net.n_input_dims, n_out,
                                            net.n_hidden_layers, net.n_neurons, net.hidden_act, out_act,
                                            y.data_ptr() + col * y.element_size(), y.shape[1], L.stream_of(x)), 'mlp_forward_f32')
        return
    L.check(L.lib().nrf_mlp_forward_ex(L.ptr(x), L.dtype_code(x.dtype), L.ptr(params_h), x.shape[0], net.n_input_dims, n_out,
                                       net.n_hidden_layers, net.n_neurons, net.hidden_act, out_act,
                                       y.data_ptr() + col * y.element_size(), L.dtype_code(y.dtype), y.shape[1],
                                       L.stream_of(x)), 'mlp_forward_ex')


def _bwd_ex(net, x, params_h, dy, col, n_out, out_act, dx, dx_accumulate, dparams):
    if params_h.dtype == torch.float32:                      # parity mode (dy must be f32)
        if dy.dtype != torch.float32:
            raise RuntimeError('nerfstyle_b200.tcnn: parity mode needs f32 output gradients')
        L.check(L.lib().nrf_mlp_backward_f32(L.ptr(x), L.dtype_code(x.dtype), L.ptr(params_h), dy.data_ptr() + col * 4, dy.shape[1],
                                             x.shape[0], net.n_input_dims, n_out, net.n_hidden_layers, net.n_neurons,
                                             net.hidden_act, out_act, L.ptr(dx), int(dx_accumulate), L.ptr(dparams),
                                             L.stream_of(x)), 'mlp_backward_f32')
        return
    L.check(L.lib().nrf_mlp_backward_ex(L.ptr(x), L.dtype_code(x.dtype), L.ptr(params_h),
                                        dy.data_ptr() + col * dy.element_size(), L.dtype_code(dy.dtype), dy.shape[1],
                                        x.shape[0], net.n_input_dims, n_out, net.n_hidden_layers, net.n_neurons,
                                        net.hidden_act, out_act, float(net.loss_scale), L.ptr(dx), L.dtype_code(x.dtype),
                                        int(dx_accumulate), L.ptr(dparams), L.stream_of(x)), 'mlp_backward_ex')


class _density_head(Function):
    """sigma = trunc_exp(density_net(enc)) as ONE kernel per direction: the exp (tcnn_nerf.py:55-69) is the output
    activation of the fused MLP (NRF_ACT_TRUNC_EXP), f32 out; replaces the f16->f32 cast + exp (+ 4 backward kernels)."""

    @staticmethod
    def forward(ctx, enc, params, net):
        L.require_cuda(enc, params)
        enc = enc.contiguous()
        params_h = half_params(params, net)
        y = torch.empty(enc.shape[0], 1, dtype=torch.float32, device=enc.device)
        with torch.cuda.device(enc.device):
            _fwd_ex(net, enc, params_h, y, 0, 1, L.ACT['trunc_exp'])
        ctx.save_for_backward(enc, params_h)
        ctx.net = net
        ctx.need = (ctx.needs_input_grad[0], ctx.needs_input_grad[1])
        return y

    @staticmethod
    def backward(ctx, g):
        enc, params_h = ctx.saved_tensors
        net = ctx.net
        g = g.contiguous()
        if g.dtype not in (torch.float16, torch.float32) or params_h.dtype == torch.float32:
            g = g.float()
        dx = torch.empty_like(enc) if ctx.need[0] else None
        dp = torch.zeros(params_h.shape, dtype=torch.float32, device=enc.device) if ctx.need[1] else None
        with torch.cuda.device(enc.device):
            _bwd_ex(net, enc, params_h, g, 0, 1, L.ACT['trunc_exp'], dx, False, dp)
        return dx, dp, None


class _color_heads(Function):
    """rgbs = cat(color2_net(color1_net(enc)), class_net(enc)) (style_nerf.py:136-141) without the cat, the f16->f32 cast
    of the compositing input and the slice / cast copies of its backward: the two output networks write column blocks
    of one f32 [B, 3+K] matrix, the backward reads column blocks of its gradient, and class_net ADDS its input gradient
    into color1_net's (REDG vector reductions) instead of a separate add over [B, 32]."""

    @staticmethod
    def forward(ctx, enc, p_class, p_c1, p_c2, class_net, color1_net, color2_net):
        L.require_cuda(enc, p_class, p_c1, p_c2)
        enc = enc.contiguous()
        B = enc.shape[0]
        K = class_net.n_output_dims
        hp = [half_params(p, n) for p, n in ((p_class, class_net), (p_c1, color1_net), (p_c2, color2_net))]
        c1 = torch.empty(B, color1_net.n_output_dims, dtype=torch.float32 if _parity else torch.float16, device=enc.device)
        rgbs = torch.empty(B, 3 + K, dtype=torch.float32, device=enc.device)
        with torch.cuda.device(enc.device):
            _fwd_ex(color1_net, enc, hp[1], c1, 0, color1_net.n_output_dims, color1_net.out_act)
            _fwd_ex(color2_net, c1, hp[2], rgbs, 0, 3, color2_net.out_act)
            _fwd_ex(class_net, enc, hp[0], rgbs, 3, K, class_net.out_act)
        ctx.save_for_backward(enc, c1, *hp)
        ctx.nets = (class_net, color1_net, color2_net)
        ctx.need = ctx.needs_input_grad[:4]
        return rgbs

    @staticmethod
    def backward(ctx, g):
        enc, c1, h_class, h_c1, h_c2 = ctx.saved_tensors
        class_net, color1_net, color2_net = ctx.nets
        K = class_net.n_output_dims
        g = g.contiguous()
        if g.dtype not in (torch.float16, torch.float32) or h_c2.dtype == torch.float32:
            g = g.float()
        dev = enc.device
        z = lambda h, need: torch.zeros(h.shape, dtype=torch.float32, device=dev) if need else None
        dp_class, dp_c1, dp_c2 = z(h_class, ctx.need[1]), z(h_c1, ctx.need[2]), z(h_c2, ctx.need[3])
        need_dx = ctx.need[0]
        dc1 = torch.empty_like(c1) if (need_dx or ctx.need[2]) else None
        dx = torch.empty_like(enc) if need_dx else None
        with torch.cuda.device(dev):
            if dc1 is not None or dp_c2 is not None:
                _bwd_ex(color2_net, c1, h_c2, g, 0, 3, color2_net.out_act, dc1, False, dp_c2)
            if dc1 is not None:
                _bwd_ex(color1_net, enc, h_c1, dc1, 0, color1_net.n_output_dims, color1_net.out_act, dx, False, dp_c1)
            if dx is not None or dp_class is not None:
                _bwd_ex(class_net, enc, h_class, g, 3, K, class_net.out_act, dx, dx is not None, dp_class)
        return dx, dp_class, dp_c1, dp_c2, None, None, None


def _field_fusable(enc_d, enc_c, density_net, class_net, color1_net, color2_net):
    """The one-launch forward (csrc/field_tc.cu) covers exactly the reference's default field: f16 32-wide encodings,
    density 32->64->1 / class 32->64->K / color1 32->64->16 (one hidden layer, linear out), color2 16->64->64->3 sigmoid."""
    def shape(n, ni, no, nh, act):
        return (n.n_input_dims, n.n_output_dims, n.n_hidden_layers, n.n_neurons, n.hidden_act, n.out_act) == (ni, no, nh, 64, _ACT['relu'], act)
    return (not _parity and _fuse_field and enc_d.is_cuda and enc_d.dtype == torch.float16 and enc_c.dtype == torch.float16
            and enc_d.shape == enc_c.shape and enc_d.shape[-1] == 32 and class_net.n_output_dims <= 16
            and shape(density_net, 32, 1, 1, _ACT['none']) and shape(class_net, 32, class_net.n_output_dims, 1, _ACT['none'])
            and shape(color1_net, 32, 16, 1, _ACT['none']) and shape(color2_net, 16, 3, 2, _ACT['sigmoid']))


_fuse_field = True


def set_field_fusion(on):
    """False: the field heads run as four launches (density_head + color_heads) -- for A/B measurements and tests."""
    global _fuse_field
    _fuse_field = bool(on)


class _field_heads(Function):
    """(rgbs, sigmas) of the whole field head in ONE forward launch (nrf_field_forward); the backward is the four
    tensor-core backward launches of _density_head / _color_heads on the saved encodings and color1 output."""

    @staticmethod
    def forward(ctx, enc_d, enc_c, p_density, p_class, p_c1, p_c2, density_net, class_net, color1_net, color2_net, ready=None):
        L.require_cuda(enc_d, enc_c, p_density, p_class, p_c1, p_c2)
        enc_d, enc_c = enc_d.contiguous(), enc_c.contiguous()
        B, K = enc_d.shape[0], class_net.n_output_dims
        hp = [half_params(p, n) for p, n in ((p_density, density_net), (p_class, class_net), (p_c1, color1_net), (p_c2, color2_net))]
        sigmas = torch.empty(B, 1, dtype=torch.float32, device=enc_d.device)
        rgbs = torch.empty(B, 3 + K, dtype=torch.float32, device=enc_d.device)
        need_bwd = any(ctx.needs_input_grad[:6])
        c1 = torch.empty(B, 16, dtype=torch.float16, device=enc_d.device) if need_bwd else None
        with torch.cuda.device(enc_d.device):
            main = torch.cuda.current_stream(enc_d.device)
            # `ready`: the encodings are still being gathered chunk by chunk on a side stream (gridencoder pipeline=True);
            # the networks of a chunk start as soon as that chunk's gather has finished
            for (r0, r1, ev) in (ready or [(0, B, None)]):
                if ev is not None:
                    main.wait_event(ev)
                L.check(L.lib().nrf_field_forward(enc_d.data_ptr() + 64 * r0, enc_c.data_ptr() + 64 * r0, L.ptr(hp[0]), L.ptr(hp[1]),
                                                  L.ptr(hp[2]), L.ptr(hp[3]), r1 - r0, K, sigmas.data_ptr() + 4 * r0,
                                                  rgbs.data_ptr() + 4 * (3 + K) * r0, 3 + K, (c1.data_ptr() + 32 * r0) if c1 is not None else None,
                                                  None, main.cuda_stream), 'field_forward')
        if need_bwd:
            ctx.save_for_backward(enc_d, enc_c, c1, *hp)
        ctx.nets = (density_net, class_net, color1_net, color2_net)
        ctx.need = ctx.needs_input_grad[:6]
        return rgbs, sigmas

    @staticmethod
    def backward(ctx, g_rgbs, g_sigmas):
        enc_d, enc_c, c1, h_d, h_class, h_c1, h_c2 = ctx.saved_tensors
        density_net, class_net, color1_net, color2_net = ctx.nets
        K = class_net.n_output_dims
        dev = enc_d.device
        z = lambda h, need: torch.zeros(h.shape, dtype=torch.float32, device=dev) if need else None      # noqa: E731
        dp_d, dp_class, dp_c1, dp_c2 = z(h_d, ctx.need[2]), z(h_class, ctx.need[3]), z(h_c1, ctx.need[4]), z(h_c2, ctx.need[5])
        dx_d = dx_c = None
        with torch.cuda.device(dev):
            if g_sigmas is not None and (ctx.need[0] or ctx.need[2]):
                g = g_sigmas.contiguous()
                g = g if g.dtype in (torch.float16, torch.float32) else g.float()
                dx_d = torch.empty_like(enc_d) if ctx.need[0] else None
                _bwd_ex(density_net, enc_d, h_d, g, 0, 1, L.ACT['trunc_exp'], dx_d, False, dp_d)
            if g_rgbs is not None and any(ctx.need[i] for i in (1, 3, 4, 5)):
                g = g_rgbs.contiguous()
                g = g if g.dtype in (torch.float16, torch.float32) else g.float()
                need_dx = ctx.need[1]
                dc1 = torch.empty_like(c1) if (need_dx or ctx.need[4]) else None
                dx_c = torch.empty_like(enc_c) if need_dx else None
                if dc1 is not None or dp_c2 is not None:
                    _bwd_ex(color2_net, c1, h_c2, g, 0, 3, color2_net.out_act, dc1, False, dp_c2)
                if dc1 is not None:
                    _bwd_ex(color1_net, enc_c, h_c1, dc1, 0, 16, color1_net.out_act, dx_c, False, dp_c1)
                if dx_c is not None or dp_class is not None:
                    _bwd_ex(class_net, enc_c, h_class, g, 3, K, class_net.out_act, dx_c, dx_c is not None, dp_class)
        return dx_d, dx_c, dp_d, dp_class, dp_c1, dp_c2, None, None, None, None, None


def field_heads(enc_d, enc_c, density_net, class_net, color1_net, color2_net):
    """(rgbs f32 [B, 3 + K], sigmas f32 [B, 1]) = (cat(color2(color1(enc_c)), class(enc_c)), trunc_exp(density(enc_d))):
    one launch when the networks have the reference's default shapes, else density_head + color_heads."""
    from .gridencoder import take_ready, wait_ready
    if _field_fusable(enc_d, enc_c, density_net, class_net, color1_net, color2_net) and enc_d.dim() == 2 \
            and L.lib().nrf_mlp_get_mode() == 0:
        ready = take_ready(enc_d)
        take_ready(enc_c)
        return _field_heads.apply(enc_d, enc_c, density_net.params, class_net.params, color1_net.params, color2_net.params,
                                  density_net, class_net, color1_net, color2_net, ready)
    wait_ready(enc_d, enc_c)
    enc_d = enc_d.reshape(-1, density_net.n_input_dims)
    enc_c = enc_c.reshape(-1, class_net.n_input_dims)
    return color_heads(enc_c, class_net, color1_net, color2_net), density_head(enc_d, density_net)


def density_head(enc, net):
    """trunc_exp(net(enc)) fused (f32 [B, 1])."""
    return _density_head.apply(enc.reshape(-1, net.n_input_dims), net.params, net)


def color_heads(enc, class_net, color1_net, color2_net):
    """cat(color2_net(color1_net(enc)), class_net(enc)) fused (f32 [B, 3 + K])."""
    return _color_heads.apply(enc.reshape(-1, class_net.n_input_dims), class_net.params, color1_net.params, color2_net.params,
                              class_net, color1_net, color2_net)


class Network(nn.Module):
    """tcnn.Network look-alike (FullyFusedMLP / CutlassMLP otypes map to the same fused kernel)."""

    def __init__(self, n_input_dims, n_output_dims, network_config, seed=1337):
        super().__init__()
        otype = network_config.get('otype', 'FullyFusedMLP')
        if otype not in ('FullyFusedMLP', 'CutlassMLP'):
            raise RuntimeError('nerfstyle_b200.tcnn.Network: unsupported otype %r' % otype)
        self.n_input_dims = int(n_input_dims)
        self.n_output_dims = int(n_output_dims)
        self.network_config = dict(network_config)
        self.seed = seed
        self.n_neurons = int(network_config.get('n_neurons', 64))
        self.n_hidden_layers = int(network_config.get('n_hidden_layers', 1))
        act = str(network_config.get('activation', 'ReLU')).lower()
        out_act = str(network_config.get('output_activation', 'None')).lower()
        if act not in ('relu', 'none') or out_act not in _ACT:
            raise RuntimeError('nerfstyle_b200.tcnn.Network: unsupported activation %r / %r' % (act, out_act))
        if self.n_neurons != 64 or not (1 <= self.n_hidden_layers <= 2) or self.n_input_dims > 64 \
                or self.n_output_dims > 16:
            raise RuntimeError('nerfstyle_b200.tcnn.Network: supported shapes are n_neurons=64, 1-2 hidden layers, '
                               'n_input_dims<=64, n_output_dims<=16')
        self.hidden_act = _ACT[act]
        self.out_act = _ACT[out_act]
        self.dtype = torch.float16
        self.loss_scale = 128.0
        self.in_pad = _pad16(self.n_input_dims)
        self.out_pad = _pad16(self.n_output_dims)
        shapes = [(self.n_neurons, self.in_pad)]
        shapes += [(self.n_neurons, self.n_neurons)] * (self.n_hidden_layers - 1)
        shapes += [(self.out_pad, self.n_neurons)]
        self.layer_shapes = shapes
        gen = torch.Generator(device='cpu')
        gen.manual_seed(int(seed))
        chunks = []
        for (o, i) in shapes:
            bound = math.sqrt(6.0 / (i + o))
            chunks.append((torch.rand(o, i, generator=gen) * 2 - 1).mul_(bound).reshape(-1))
        self.params = nn.Parameter(torch.cat(chunks))

    def layer_views(self, flat=None):
        """The per-layer [out_pad, in_pad] views of a flat parameter vector (default: self.params)."""
        flat = self.params if flat is None else flat
        views, o = [], 0
        for (r, c) in self.layer_shapes:
            views.append(flat[o:o + r * c].view(r, c))
            o += r * c
        return views

    def forward(self, x):
        cfg = (self.n_input_dims, self.n_output_dims, self.n_hidden_layers, self.n_neurons, self.hidden_act,
               self.out_act, self.loss_scale)
        lead = x.shape[:-1]
        y = _mlp_function.apply(x.reshape(-1, self.n_input_dims), self.params, cfg, self)
        return y.view(*lead, self.n_output_dims)

    def extra_repr(self):
        return 'n_input_dims=%d, n_output_dims=%d, n_hidden_layers=%d, n_neurons=%d' % (
            self.n_input_dims, self.n_output_dims, self.n_hidden_layers, self.n_neurons)


class Encoding(nn.Module):
    """tcnn.Encoding look-alike for {'otype': 'SphericalHarmonics', 'degree': d <= 4} (networks/style_nerf.py:33-42,
    networks/tcnn_nerf.py:87-95; only built when the model uses view directions).  Inputs in [0,1]^3 (the callers map
    directions with (d + 1) / 2), output `dtype` [B, d*d].  The encoding has no parameters (`params` is empty, as in
    tinycudann) and is forward-only: ray directions never carry gradients on this path."""

    def __init__(self, n_input_dims, encoding_config, seed=1337, dtype=None):
        super().__init__()
        otype = encoding_config.get('otype')
        if otype != 'SphericalHarmonics':
            raise NotImplementedError('nerfstyle_b200.tcnn.Encoding: only SphericalHarmonics is implemented (got %r)' % (otype,))
        if int(n_input_dims) != 3:
            raise RuntimeError('nerfstyle_b200.tcnn.Encoding(SphericalHarmonics) needs n_input_dims = 3')
        self.n_input_dims = 3
        self.degree = int(encoding_config.get('degree', 4))
        if not 1 <= self.degree <= 4:
            raise RuntimeError('nerfstyle_b200.tcnn.Encoding(SphericalHarmonics): degree 1..4 supported')
        self.n_output_dims = self.degree * self.degree
        self.encoding_config = dict(encoding_config)
        self.seed = seed
        self.dtype = torch.float16 if dtype is None else dtype
        self.params = nn.Parameter(torch.zeros(0), requires_grad=False)

    def forward(self, x):
        L.require_cuda(x)
        if x.requires_grad:
            raise NotImplementedError('nerfstyle_b200.tcnn.Encoding(SphericalHarmonics) is forward-only')
        lead = x.shape[:-1]
        xi = x.reshape(-1, 3).float().contiguous()
        out = torch.empty(xi.shape[0], self.n_output_dims, dtype=self.dtype, device=x.device)
        with torch.cuda.device(x.device):
            L.check(L.lib().nrf_sh_encode_forward(L.ptr(xi), xi.shape[0], self.degree, L.ptr(out), L.dtype_code(self.dtype),
                                                  L.stream_of(xi)), 'sh_encode_forward')
        return out.view(*lead, self.n_output_dims)
