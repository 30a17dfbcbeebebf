"""torch-CPU restatement of the model glue and the tcnn-style MLP -- TEST INFRASTRUCTURE ONLY.

Follows networks/style_nerf.py:120-142 (StyleTCNerf._forward), networks/tcnn_nerf.py:55-69 (trunc_exp),
common.py:276-288 (BBox.normalize), gridencoder/grid.py:173-191 (GridEncoder.forward remap) and
renderer.py:196-235 (render_train) of the reference.  The MLP is the oracle *definition* of SURVEY.md 8c
(tiny-cuda-nn is un-vendored and unpinned): y = act_out(W_n relu(... relu(W_1 x))), no bias, weights
[out,in]; in "half" mode the weights, the layer inputs and the hidden activations are rounded to fp16
exactly where the CUDA kernel rounds them, products accumulate in fp32.
"""
import numpy as np
import torch
from torch.autograd import Function

from . import ops


def _round_half(t):
    """fp16 rounding with a straight-through gradient."""
    return t + (t.half().float() - t).detach()


# ------------------------------------------------------------------------------------------ autograd wrappers
class GridEncodeFn(Function):
    """grid.py:19-97 on the C oracle (float tables, or fp16 tables/grads when half=True)."""

    @staticmethod
    def forward(ctx, inputs, embeddings, offsets, per_level_scale, base_resolution, gridtype, align_corners, style,
                half):
        inp = inputs.detach().float().contiguous()
        emb = embeddings.detach()
        out, _ = ops.grid_encode_forward(inp.numpy(), emb.half().numpy() if half else emb.float().numpy(),
                                         offsets.numpy(), per_level_scale, base_resolution, False, gridtype,
                                         align_corners, style, half=half)
        ctx.save_for_backward(inp, offsets)
        ctx.cfg = (per_level_scale, base_resolution, gridtype, align_corners, style, half, emb.shape)
        return torch.from_numpy(out)

    @staticmethod
    def backward(ctx, grad):
        inp, offsets = ctx.saved_tensors
        pls, H, gridtype, ac, style, half, shape = ctx.cfg
        g = grad.contiguous()
        # fp16 mode: the incoming grads are fp16 (the encoder output dtype); the product path accumulates them in
        # fp32 (deviation from the reference's lossy __half2 atomics), so the oracle value is the exact sum
        g = g.half().float() if half else g.float()
        ge = ops.grid_encode_backward(g.numpy(), inp.numpy(), offsets.numpy(), shape[0], shape[1], pls, H, gridtype, ac,
                                      style, half=False)
        return None, torch.from_numpy(ge).float(), None, None, None, None, None, None, None


class CompositeTrainFn(Function):
    """raymarching.py:291-350 on the C oracle."""

    @staticmethod
    def forward(ctx, sigmas, rgbs, deltas, rays, T_thresh, is_ndc):
        s = sigmas.detach().float().contiguous().view(-1)
        r = rgbs.detach().float().contiguous()
        ws, depth, image = ops.composite_rays_train_forward(s.numpy(), r.numpy(), deltas.numpy(), rays.numpy(), T_thresh,
                                                            is_ndc)
        ws, depth, image = torch.from_numpy(ws), torch.from_numpy(depth), torch.from_numpy(image)
        ctx.save_for_backward(s, r, deltas, rays, ws, image)
        ctx.cfg = (T_thresh, is_ndc, sigmas.shape)
        return ws, depth, image

    @staticmethod
    def backward(ctx, g_ws, g_depth, g_image):
        s, r, deltas, rays, ws, image = ctx.saved_tensors
        T_thresh, is_ndc, sshape = ctx.cfg
        gs, gr = ops.composite_rays_train_backward(g_ws.contiguous().numpy(), g_image.contiguous().numpy(), s.numpy(),
                                                   r.numpy(), deltas.numpy(), rays.numpy(), ws.numpy(), image.numpy(),
                                                   T_thresh, is_ndc)
        return torch.from_numpy(gs).view(sshape), torch.from_numpy(gr), None, None, None, None


class TruncExpFn(Function):
    """tcnn_nerf.py:55-69"""

    @staticmethod
    def forward(ctx, x):
        x = x.float()
        ctx.save_for_backward(x)
        return torch.exp(x)

    @staticmethod
    def backward(ctx, g):
        x = ctx.saved_tensors[0]
        return g * torch.exp(x.clamp(-15, 15))


class HalfCast(Function):
    """An fp16 tensor boundary: the value is rounded to fp16 in forward and so is its gradient in backward (autograd
    hands an fp16 gradient to the producer of an fp16 tensor)."""

    @staticmethod
    def forward(ctx, x):
        return x.half().float()

    @staticmethod
    def backward(ctx, g):
        return g.half().float()


# ------------------------------------------------------------------------------------------ MLP
def mlp_layer_shapes(n_in, n_out, n_hidden, width=64):
    pad = lambda n: (n + 15) // 16 * 16  # noqa: E731
    return [(width, pad(n_in))] + [(width, width)] * (n_hidden - 1) + [(pad(n_out), width)]


def mlp_split(params, n_in, n_out, n_hidden, width=64):
    views, o = [], 0
    for (r, c) in mlp_layer_shapes(n_in, n_out, n_hidden, width):
        views.append(params[o:o + r * c].view(r, c))
        o += r * c
    return views


_ACTS = {
    'none': lambda z: z,
    'relu': torch.relu,
    'sigmoid': torch.sigmoid,
    'exponential': torch.exp,
}


def mlp_forward(x, params, n_in, n_out, n_hidden, hidden_act='relu', out_act='none', half=True, width=64, x_half=False):
    """Oracle MLP.  x [B, n_in] (any float dtype), params flat fp32 (tcnn layout).  Returns fp32 [B, n_out] holding
    fp16-representable values when half=True (the kernel's output dtype is fp16, hence so is the gradient it
    receives).  x_half: the input tensor is fp16 (so the input gradient is rounded to fp16 too)."""
    Ws = mlp_split(params, n_in, n_out, n_hidden, width)
    h = x.float()
    if x_half:
        h = HalfCast.apply(h)
    in_pad = Ws[0].shape[1]
    if in_pad > n_in:
        h = torch.nn.functional.pad(h, (0, in_pad - n_in))
    rnd = _round_half if half else (lambda t: t)
    h = rnd(h)
    for W in Ws[:-1]:
        h = rnd(_ACTS[hidden_act](h @ rnd(W).t()))
    z = h @ rnd(Ws[-1]).t()
    y = _ACTS[out_act](z)[:, :n_out]
    return HalfCast.apply(y) if half else y


# ------------------------------------------------------------------------------------------ the field
class OracleField:
    """StyleTCNerf(use_dir=False) on CPU: two hash grids + density / class / color1 / color2 MLPs."""

    def __init__(self, bound=2.0, n_classes=8, num_levels=16, level_dim=2, base_resolution=16, log2_hashmap_size=19,
                 max_res_coeff=1024, half=False, seed=0, table_std=1e-4, mlp_half=True):
        self.bound = float(bound)
        self.K = n_classes
        self.half = half
        # mlp_half=False: the fp32 "parity mode" definition of the MLP -- no fp16 rounding anywhere
        self.mlp_half = mlp_half
        # BBox(-bound, bound): size = 2*bound (style_nerf.py:28, tcnn_nerf.py:20-22)
        self.bbox_min = torch.full((3,), -self.bound)
        self.bbox_size = torch.full((3,), 2 * self.bound)
        max_res = max_res_coeff * 2 * self.bound
        pls = np.exp2(np.log2(max_res / base_resolution) / (num_levels - 1))
        offs, _ = ops.grid_offsets(3, num_levels, level_dim, pls, base_resolution, log2_hashmap_size, None, True)
        self.offsets = torch.from_numpy(offs)
        self.per_level_scale = pls
        self.base_resolution = base_resolution
        g = torch.Generator().manual_seed(seed)
        n_rows = int(offs[-1])
        self.params = {}
        self.params['x_density_embedder.embeddings'] = ((torch.rand(n_rows, level_dim, generator=g) * 2 - 1) * table_std)
        self.params['x_color_embedder.embeddings'] = ((torch.rand(n_rows, level_dim, generator=g) * 2 - 1) * table_std)
        enc_dim = num_levels * level_dim
        self.nets = {'density_net': (enc_dim, 1, 1, 'none'), 'class_net': (enc_dim, n_classes, 1, 'none'),
                     'color1_net': (enc_dim, 16, 1, 'none'), 'color2_net': (16, 3, 2, 'sigmoid')}
        for name, (ni, no, nh, _) in self.nets.items():
            chunks = []
            for (o, i) in mlp_layer_shapes(ni, no, nh):
                b = (6.0 / (i + o)) ** 0.5
                chunks.append(((torch.rand(o, i, generator=g) * 2 - 1) * b).reshape(-1))
            self.params[name + '.params'] = torch.cat(chunks)
        for p in self.params.values():
            p.requires_grad_(True)

    def encode(self, which, pts01):
        # GridEncoder.forward with the default bound=1 (grid.py:177): [0,1] -> [0.5,1]
        inp = (pts01 + 1) / 2
        return GridEncodeFn.apply(inp, self.params[which + '.embeddings'], self.offsets, self.per_level_scale,
                                  self.base_resolution, 0, True, 0, self.half)

    def net(self, name, x):
        ni, no, nh, oact = self.nets[name]
        # color2_net always consumes an fp16 tensor (color1's output); the others consume the encoder output, which is
        # fp16 only under autocast
        if not self.mlp_half:
            return mlp_forward(x, self.params[name + '.params'], ni, no, nh, 'relu', oact, half=False, x_half=self.half and name != 'color2_net')
        return mlp_forward(x, self.params[name + '.params'], ni, no, nh, 'relu', oact, half=True,
                           x_half=(self.half or name == 'color2_net'))

    def forward(self, pts, dirs=None):
        pts01 = (pts - self.bbox_min) / self.bbox_size          # common.py:288
        sigmas = TruncExpFn.apply(self.net('density_net', self.encode('x_density_embedder', pts01)))
        if dirs is None:
            return sigmas
        xc = self.encode('x_color_embedder', pts01)
        classes = self.net('class_net', xc)
        rgb = self.net('color2_net', self.net('color1_net', xc))
        return torch.cat((rgb, classes), dim=1), sigmas


def render_train(field, rays_o, rays_d, bitfield, cascade, grid_size, bound, min_near=0.2, max_steps=1024,
                 T_thresh=1e-4, density_scale=1.0):
    """renderer.py:196-235 on the oracle ops (force_all_rays=True, align=128, dt_gamma=0)."""
    aabb = np.array([-bound, -bound, -bound, bound, bound, bound], np.float32)
    nears, fars = ops.near_far_from_aabb(rays_o, rays_d, aabb, min_near)
    counter = np.zeros(2, np.int32)
    xyzs, dirs, deltas, rays = ops.march_rays_train(rays_o, rays_d, None, bound, bitfield, cascade, grid_size, nears, fars,
                                                    counter, -1, True, 128, True, 0., max_steps, False)
    rgbs, sigmas = field.forward(torch.from_numpy(xyzs.copy()), torch.from_numpy(dirs.copy()))
    sigmas = sigmas * density_scale
    ws, depth, image = CompositeTrainFn.apply(sigmas, rgbs.float(), torch.from_numpy(deltas.copy()),
                                              torch.from_numpy(rays.copy()), T_thresh, False)
    classes = image[:, 3:]
    rgb = image[:, :3] + (1 - ws).unsqueeze(-1)
    nears_t, fars_t = torch.from_numpy(nears), torch.from_numpy(fars)
    depth_n = torch.clamp(depth - nears_t, min=0) / (fars_t - nears_t)
    return {'rgb': rgb, 'depth': depth_n, 'classes': classes, 'weights_sum': ws, 'image': image, 'rays': rays,
            'counter': counter, 'xyzs': xyzs, 'deltas': deltas, 'sigmas': sigmas, 'rgbs': rgbs}


def train_step_loss(out, target_rgb, target_cls=None, class_lambda=0.001):
    """trainers/base.py:251-304: MSE + lambda * cross-entropy on the class channels."""
    loss = torch.mean((out['rgb'] - target_rgb) ** 2)
    if target_cls is not None:
        loss = loss + class_lambda * torch.nn.functional.cross_entropy(out['classes'], target_cls)
    return loss


def sh_encode(dirs01, degree=4):
    """tcnn 'SphericalHarmonics' encoding (tiny-cuda-nn is un-vendored and unpinned, README.md:25-26; call sites
    networks/style_nerf.py:33-42, networks/tcnn_nerf.py:87-95).  Restates the published real spherical-harmonics basis:
    inputs in [0,1]^3 -> [-1,1]^3, Y_lm as polynomials in (x, y, z), l < degree <= 4.  Pinned by the orthonormality of the
    basis over the unit sphere (tests/test_oracle_kat.py), not by the reference (parity unpinned)."""
    import numpy as np
    d = np.asarray(dirs01, dtype=np.float64) * 2.0 - 1.0
    x, y, z = d[:, 0], d[:, 1], d[:, 2]
    xy, xz, yz, x2, y2, z2 = x * y, x * z, y * z, x * x, y * y, z * z
    o = [np.full_like(x, 0.28209479177387814),
         -0.48860251190291987 * y, 0.48860251190291987 * z, -0.48860251190291987 * x,
         1.0925484305920792 * xy, -1.0925484305920792 * yz, 0.94617469575755997 * z2 - 0.31539156525251999,
         -1.0925484305920792 * xz, 0.54627421529603959 * x2 - 0.54627421529603959 * y2,
         0.59004358992664352 * y * (-3.0 * x2 + y2), 2.8906114426405538 * xy * z, 0.45704579946446572 * y * (1.0 - 5.0 * z2),
         0.3731763325901154 * z * (5.0 * z2 - 3.0), 0.45704579946446572 * x * (1.0 - 5.0 * z2),
         1.4453057213202769 * z * (x2 - y2), 0.59004358992664352 * x * (-x2 + 3.0 * y2)]
    return np.stack(o[:degree * degree], axis=1)
