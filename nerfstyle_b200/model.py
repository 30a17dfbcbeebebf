"""Host-side mirror of the reference's field and renderer glue, on the drop-in ops.

The reference's own `networks/style_nerf.py`, `networks/tcnn_nerf.py` and `renderer.py` run unchanged on
nerfstyle_b200.dropin (they only need `raymarching`, `gridencoder`, `tinycudann`), but they also import the
reference's config / utils / nerf_lib stack (dacite, simple_parsing, torch_ema, imageio ... none of which is
needed by the hot path or present on the GPU box).  This module restates the *call pattern* of those files
without that stack, with identical parameter names and state-dict keys, so bench.py / tests can drive the
path exactly the way the reference does:

  StyleTCNerf   networks/style_nerf.py:12-159   (use_dir=False default path, trainers/base.py:149-151)
  trunc_exp     networks/tcnn_nerf.py:55-69
  Renderer      renderer.py:19-293              (update_state, render_train, render_test)
"""
import itertools
from math import ceil, log2

import numpy as np
import torch
import torch.nn as nn
from torch.autograd import Function
from torch.amp import custom_bwd, custom_fwd

from . import raymarching
from . import tcnn
from .gridencoder import GridEncoder, grid_encode_dual, same_geometry
from .optim import current_half_copy

STEP_CTR_SIZE = 16


class _trunc_exp(Function):
    """tcnn_nerf.py:55-69"""

    @staticmethod
    @custom_fwd(device_type='cuda', cast_inputs=torch.float32)
    def forward(ctx, x):
        x = x.float()      # the reference only runs under autocast, where cast_inputs makes this fp32
        ctx.save_for_backward(x)
        return torch.exp(x)

    @staticmethod
    @custom_bwd(device_type='cuda')
    def backward(ctx, g):
        x = ctx.saved_tensors[0]
        return g * torch.exp(x.clamp(-15, 15))


trunc_exp = _trunc_exp.apply


def get_grid_encoder(n_lvls=16, n_feats_per_lvl=2, hashmap_size=19, min_res=16, max_res_coeff=1024, max_bound=4.0):
    """tcnn_nerf.py:14-35 with cfgs/network/default.yaml values as defaults."""
    max_res = max_res_coeff * max_bound
    per_lvl_scale = np.exp2(np.log2(max_res / min_res) / (n_lvls - 1))
    return GridEncoder(input_dim=3, num_levels=n_lvls, level_dim=n_feats_per_lvl, per_level_scale=per_lvl_scale,
                       base_resolution=min_res, log2_hashmap_size=hashmap_size, gridtype='hash', align_corners=True)


def _net(n_in, n_out, n_hidden, out_act, seed):
    return tcnn.Network(n_input_dims=n_in, n_output_dims=n_out, network_config={
        'otype': 'FullyFusedMLP', 'activation': 'ReLU', 'output_activation': out_act, 'n_neurons': 64,
        'n_hidden_layers': n_hidden}, seed=seed)


class StyleTCNerf(nn.Module):
    """networks/style_nerf.py:12-159, use_dir=False."""

    def __init__(self, bbox_min, bbox_max, class_dim, density_hidden_layers=1, rgb_hidden_layers=2, network_seed=80000,
                 fused_heads=True, **enc_kwargs):
        super().__init__()
        # fused_heads: same math as the reference's op-by-op glue below, with trunc_exp / cat / casts folded into the
        # MLP kernels (tcnn.density_head / tcnn.color_heads); False runs the reference's exact op sequence
        self.fused_heads = fused_heads
        self._dual = None
        # not persistent: the reference's bounding box is a plain BBox attribute, so its state_dict has no such keys
        self.register_buffer('bbox_min', torch.as_tensor(bbox_min, dtype=torch.float32), persistent=False)
        self.register_buffer('bbox_size', torch.as_tensor(bbox_max, dtype=torch.float32) - self.bbox_min, persistent=False)
        self.class_dim = class_dim
        self.use_dir = False
        max_bound = float(torch.max(self.bbox_size).item())
        self.x_density_embedder = get_grid_encoder(max_bound=max_bound, **enc_kwargs)
        self.x_color_embedder = get_grid_encoder(max_bound=max_bound, **enc_kwargs)
        self.density_net = _net(self.x_density_embedder.n_output_dims, 1, density_hidden_layers, 'None', network_seed)
        self.color1_net = _net(self.x_color_embedder.n_output_dims, 16, density_hidden_layers, 'None', network_seed)
        self.color2_net = _net(self.color1_net.n_output_dims, 3, rgb_hidden_layers, 'Sigmoid', network_seed)
        self.class_net = _net(self.x_color_embedder.n_output_dims, class_dim, density_hidden_layers, 'None',
                              network_seed)

    @property
    def device(self):
        return self.bbox_min.device

    def _apply(self, fn, *args, **kwargs):
        out = super()._apply(fn, *args, **kwargs)
        self._dual = None                  # .to() / .cuda() moved the buffers: the cached kernel-side transform is rebuilt
        return out

    def state_dict(self, *args, **kwargs):
        """A sharded fused optimizer keeps the fp32 master tables current only inside each rank's shard: gather them first
        (a collective when world_size > 1 -- every rank must call state_dict())."""
        from .optim import sync_for_checkpoint
        sync_for_checkpoint(self)
        return super().state_dict(*args, **kwargs)

    def _forward(self, pts, dirs=None):
        if self.fused_heads and pts.is_cuda and dirs is not None:
            if self._dual is None:
                self._dual = bool(same_geometry(self.x_density_embedder, self.x_color_embedder))
                self._xform = torch.cat([self.bbox_min, self.bbox_size, self.bbox_min.new_ones(1)]).contiguous()
            if self._dual and pts.dtype == torch.float32:
                # one index computation for both hash tables; BBox.normalize + the encoder's input map run in the kernel
                # pipeline=True: the gather runs chunk by chunk on a side stream and field_heads starts the networks of a
                # chunk as soon as its encodings exist (the two kernels co-reside on the SMs: the gather holds no shared
                # memory / TMEM)
                x_embedded, x_color_embedded = grid_encode_dual(pts, self.x_density_embedder, self.x_color_embedder,
                                                                xform=self._xform, pipeline=pts.dim() == 2)
                return tcnn.field_heads(x_embedded, x_color_embedded, self.density_net, self.class_net, self.color1_net,
                                        self.color2_net)
        pts = (pts - self.bbox_min) / self.bbox_size            # BBox.normalize, common.py:288
        if self.fused_heads and pts.is_cuda:
            if dirs is None:
                return tcnn.density_head(self.x_density_embedder(pts), self.density_net)
            if self._dual:
                x_embedded, x_color_embedded = grid_encode_dual(pts, self.x_density_embedder, self.x_color_embedder)
            else:
                x_embedded, x_color_embedded = self.x_density_embedder(pts), self.x_color_embedder(pts)
            return tcnn.field_heads(x_embedded, x_color_embedded, self.density_net, self.class_net, self.color1_net,
                                    self.color2_net)
        x_embedded = self.x_density_embedder(pts)
        density_output = self.density_net(x_embedded)
        sigmas = trunc_exp(density_output)
        if dirs is None:
            return sigmas
        x_color_embedded = self.x_color_embedder(pts)
        classes = self.class_net(x_color_embedded)
        color1_output = self.color1_net(x_color_embedded)
        rgbs = self.color2_net(color1_output)
        rgbs = torch.cat((rgbs, classes), dim=1)
        return rgbs, sigmas

    def forward(self, pts, dirs=None, bsize=1 << 24):
        """style_nerf.py:144-159: batches of `bsize` points when N >= bsize.  The reference chunks at 10^6 points to fit
        a 24 GB card; every op is point-wise, so the chunk size does not change any value -- on a 180 GB B200 the
        default is 2^24 points (one launch per op for any realistic ray batch; pass bsize=1000000 for the reference's)."""
        N = len(pts)
        if N < bsize:
            return self._forward(pts, dirs)
        sigmas = torch.empty((N, 1), device=self.device)
        rgbs = torch.empty((N, 3 + self.class_dim), device=self.device) if dirs is not None else None
        for s in range(0, N, bsize):
            e = min(N, s + bsize)
            if dirs is None:
                sigmas[s:e] = self._forward(pts[s:e])
            else:
                r, sg = self._forward(pts[s:e], dirs[s:e])
                rgbs[s:e] = r
                sigmas[s:e] = sg
        return sigmas if dirs is None else (rgbs, sigmas)


class Renderer(nn.Module):
    """renderer.py:19-293.  Config values default to cfgs/renderer/default.yaml + llff.yaml."""

    def __init__(self, model, bound, raymarch_channels=3, grid_size=128, update_iter=16, update_thres=256,
                 density_thresh=10., density_decay=0.95, density_scale=1., min_near=0.2, t_thresh=1e-4, max_steps=1024,
                 grid_bsize=None, fused_occupancy=True):
        super().__init__()
        self.model = model
        self.bound = bound
        self.raymarch_channels = raymarch_channels
        self.grid_size, self.update_iter, self.update_thres = grid_size, update_iter, update_thres
        self.density_thresh, self.density_decay, self.density_scale = density_thresh, density_decay, density_scale
        self.min_near, self.t_thresh, self.max_steps, self.grid_bsize = min_near, t_thresh, max_steps, grid_bsize
        self.update_occ = True
        # fused_occupancy: update_state runs as device passes with no host read-back (csrc/occupancy.cu); False = the
        # reference's op sequence (update_state_reference)
        self.fused_occupancy = fused_occupancy
        self._occ = None
        self.cascade = 1 + ceil(log2(bound))
        self.register_buffer('aabb', torch.tensor([-bound, -bound, -bound, bound, bound, bound], dtype=torch.float32),
                             persistent=False)
        self.register_buffer('density_grid', torch.zeros((self.cascade, grid_size ** 3)), persistent=False)
        self.register_buffer('density_bitfield', torch.zeros((self.cascade * grid_size ** 3 // 8,), dtype=torch.uint8),
                             persistent=False)
        self.register_buffer('step_counter', torch.zeros((STEP_CTR_SIZE, 2), dtype=torch.int32), persistent=False)
        self.local_step = 0
        self.mean_count = 0
        self.mean_density = 0

    @property
    def device(self):
        return self.aabb.device

    # mean_density / mean_count (renderer.py:187,194) are host numbers in the reference, read back inside update_state.
    # The fused update leaves them on the device; they are fetched (one sync) only when somebody asks.
    @property
    def mean_density(self):
        v = self._mean_density
        if torch.is_tensor(v):
            v = self._mean_density = v.item()
        return v

    @mean_density.setter
    def mean_density(self, v):
        self._mean_density = v

    @property
    def mean_count(self):
        v = self._mean_count
        if isinstance(v, tuple):
            v = self._mean_count = int(v[0].item() / v[1])
        return v

    @mean_count.setter
    def mean_count(self, v):
        self._mean_count = v

    def state_dict(self, *args, **kwargs):
        """renderer.py:78-91 (same keys; intr / precrop_frac belong to the caller's camera set-up)."""
        sd = {'model': self.model.state_dict()}
        for k in ['raymarch_channels', 'bound', 'density_grid', 'density_bitfield', 'step_counter', 'local_step',
                  'mean_count', 'mean_density']:
            v = getattr(self, k)
            sd[k] = v.detach() if torch.is_tensor(v) else v
        return sd

    def load_state_dict(self, state_dict, *args, **kwargs):
        """renderer.py:93-107"""
        for k in ['raymarch_channels', 'bound']:
            if getattr(self, k) != state_dict[k]:
                raise RuntimeError('Values do not match when loading key "{}"'.format(k))
        self.model.load_state_dict(state_dict['model'])
        for k in ['density_grid', 'density_bitfield', 'step_counter']:
            getattr(self, k).copy_(state_dict[k].to(self.device))
        for k in ['local_step', 'mean_count', 'mean_density']:
            setattr(self, k, state_dict[k])

    def _compute_occ_sigmas(self, xyzs, cas):
        """renderer.py:120-136"""
        bound = min(2 ** cas, self.bound)
        half_grid_size = bound / self.grid_size
        cas_xyzs = xyzs * (bound - half_grid_size)
        cas_xyzs += (torch.rand_like(cas_xyzs) * 2 - 1) * half_grid_size
        sigmas = self.model(cas_xyzs).reshape(-1).detach() * self.density_scale
        return sigmas

    def update_state(self):
        """renderer.py:138-194."""
        if self.fused_occupancy and self.density_grid.is_cuda:
            return self.update_state_fused()
        return self.update_state_reference()

    def _occ_buffers(self):
        L = raymarching.L
        if self._occ is None or self._occ['dev'] != self.device:
            sh = []
            for cas in range(self.cascade):                      # renderer.py:124-127
                bound = min(2 ** cas, self.bound)
                half_grid_size = bound / self.grid_size
                sh += [bound - half_grid_size, half_grid_size]
            self._occ = {'dev': self.device,
                         'scale_hgs': torch.tensor(sh, dtype=torch.float32, device=self.device),
                         'state': torch.zeros(2, dtype=torch.float32, device=self.device),
                         'scratch': torch.empty(int(L.lib().nrf_occ_scratch_bytes()), dtype=torch.uint8, device=self.device)}
        return self._occ

    @torch.no_grad()
    def update_state_fused(self, noise=None, rnd_cells=None, pick=None):
        """renderer.py:138-194 as device passes: the sample points of every cascade come out of one kernel in Morton order
        (so the density query's output IS tmp_grid -- no meshgrid / cat / morton / index_put), ONE density query covers all
        cascades, the decay / max update, the mean and the packbits threshold stay on the device (the reference reads the
        mean density, the occupied-cell lists and the mean sample count back to the host every update).  The random inputs
        (torch.rand_like :129, torch.randint :160,:164) can be passed in for testing."""
        L = raymarching.L
        lib, dev = L.lib(), self.device
        H, C = self.grid_size, self.cascade
        H3 = H ** 3
        ob = self._occ_buffers()
        grid = self.density_grid
        assert grid.is_contiguous() and grid.dtype == torch.float32
        with torch.cuda.device(dev):
            st = torch.cuda.current_stream(dev).cuda_stream
            if self.local_step < self.update_thres:
                if noise is None:
                    noise = torch.rand(C, H3, 3, device=dev)
                pts = torch.empty(C, H3, 3, dtype=torch.float32, device=dev)
                L.check(lib.nrf_occ_points_full(pts.data_ptr(), L.ptr(noise), H, C, ob['scale_hgs'].data_ptr(), st), 'occ_points_full')
                values = self.model(pts.view(-1, 3)).reshape(-1).detach().float().contiguous()
                tmp_scale = float(self.density_scale)
            else:
                N = H3 // 4
                if rnd_cells is None:
                    rnd_cells = torch.randint(0, H, (C, N, 3), device=dev, dtype=torch.int32)
                if pick is None:
                    pick = torch.rand(C, N, device=dev)
                if noise is None:
                    noise = torch.rand(C, 2 * N, 3, device=dev)
                flags = torch.empty(C, H3, dtype=torch.int32, device=dev)
                occ_list = torch.empty(C, H3, dtype=torch.int32, device=dev)
                occ_count = torch.empty(C, dtype=torch.int32, device=dev)
                L.check(lib.nrf_occ_flags(grid.data_ptr(), H, C, flags.data_ptr(), st), 'occ_flags')
                scratch = L.scratch(dev, lib.nrf_march_scratch_bytes(H3))
                for cas in range(C):                 # the device-side torch.nonzero(density_grid[cas] > 0) of :163
                    L.check(lib.nrf_compact_alive(flags[cas].data_ptr(), H3, occ_list[cas].data_ptr(),
                                                  occ_count[cas:cas + 1].data_ptr(), scratch.data_ptr(), st), 'compact_alive')
                pts = torch.empty(C, 2 * N, 3, dtype=torch.float32, device=dev)
                indices = torch.empty(C, 2 * N, dtype=torch.int32, device=dev)
                L.check(lib.nrf_occ_points_sparse(pts.data_ptr(), indices.data_ptr(), L.ptr(noise), rnd_cells.contiguous().data_ptr(),
                                                  pick.contiguous().data_ptr(), occ_list.data_ptr(), occ_count.data_ptr(), N, H, C,
                                                  ob['scale_hgs'].data_ptr(), st), 'occ_points_sparse')
                sig = self.model(pts.view(-1, 3)).reshape(-1).detach().float().contiguous()
                values = torch.full((C, H3), -1.0, dtype=torch.float32, device=dev)
                L.check(lib.nrf_occ_scatter_max(values.data_ptr(), indices.data_ptr(), sig.data_ptr(), float(self.density_scale),
                                                2 * N, H, C, st), 'occ_scatter_max')
                tmp_scale = 1.0
                self._occ_last = (indices, pts)
            L.check(lib.nrf_occ_update(grid.data_ptr(), values.data_ptr(), tmp_scale, float(self.density_decay), C * H3,
                                       float(self.density_thresh), ob['state'].data_ptr(), ob['scratch'].data_ptr(), st), 'occ_update')
            L.check(lib.nrf_packbits_dev(grid.data_ptr(), C * H3 // 8, ob['state'].data_ptr() + 4, self.density_bitfield.data_ptr(), st),
                    'packbits_dev')
        self._mean_density = ob['state'][0].clone()                       # fetched lazily (property)
        total_step = min(STEP_CTR_SIZE, self.update_iter)
        self._mean_count = (self.step_counter[:total_step, 0].sum(), total_step)

    @torch.no_grad()
    def update_state_reference(self):
        """renderer.py:138-194, op for op."""
        tmp_grid = -torch.ones_like(self.density_grid)
        if self.local_step < self.update_thres:
            bsize = self.grid_bsize if self.grid_bsize is not None else self.grid_size
            X, Y, Z = [torch.arange(self.grid_size, dtype=torch.int32, device=self.device).split(bsize) for _ in range(3)]
            for (xs, ys, zs) in itertools.product(X, Y, Z):
                xx, yy, zz = torch.meshgrid(xs, ys, zs, indexing='ij')
                coords = torch.cat([xx.reshape(-1, 1), yy.reshape(-1, 1), zz.reshape(-1, 1)], dim=-1)
                indices = raymarching.morton3D(coords).long()
                xyzs = 2 * coords.float() / (self.grid_size - 1) - 1
                for cas in range(self.cascade):
                    tmp_grid[cas, indices] = self._compute_occ_sigmas(xyzs, cas)
        else:
            N = self.grid_size ** 3 // 4
            for cas in range(self.cascade):
                coords = torch.randint(0, self.grid_size, (N, 3), device=self.device)
                indices = raymarching.morton3D(coords).long()
                occ_indices = torch.nonzero(self.density_grid[cas] > 0).squeeze(-1)
                rand_mask = torch.randint(0, occ_indices.shape[0], [N], dtype=torch.long, device=self.device)
                occ_indices = occ_indices[rand_mask]
                occ_coords = raymarching.morton3D_invert(occ_indices)
                indices = torch.cat([indices, occ_indices], dim=0)
                coords = torch.cat([coords, occ_coords], dim=0)
                xyzs = 2 * coords.float() / (self.grid_size - 1) - 1
                tmp_grid[cas, indices] = self._compute_occ_sigmas(xyzs, cas)
        valid_mask = (self.density_grid >= 0) & (tmp_grid >= 0)
        self.density_grid[valid_mask] = torch.maximum(self.density_grid[valid_mask] * self.density_decay,
                                                      tmp_grid[valid_mask])
        self.mean_density = torch.mean(self.density_grid.clamp(min=0)).item()
        density_thresh = min(self.mean_density, self.density_thresh)
        self.density_bitfield = raymarching.packbits(self.density_grid, density_thresh, self.density_bitfield)
        total_step = min(STEP_CTR_SIZE, self.update_iter)
        self.mean_count = int(self.step_counter[:total_step, 0].sum().item() / total_step)

    def render_train_raw(self, rays_o, rays_d, **kwargs):
        """renderer.py:196-228: everything up to the compositing outputs.  Returns (weights_sum [N], depth [N] (relative to
        the near plane, un-normalised), image [N, 3 + K] (rgb without background, class logits), nears, fars)."""
        if self.update_occ and (self.local_step % self.update_iter == 0):
            self.update_state()
        nears, fars = raymarching.near_far_from_aabb(rays_o, rays_d, self.aabb, self.min_near)
        if self.update_occ:
            counter = self.step_counter[self.local_step % STEP_CTR_SIZE]
            counter.zero_()
            self.local_step += 1
        else:
            counter = torch.zeros(2).to(self.step_counter)
        xyzs, dirs, deltas, rays_info = raymarching.march_rays_train(
            rays_o, rays_d, None, self.bound, self.density_bitfield, self.cascade, self.grid_size, nears, fars, counter,
            self._mean_count if isinstance(self._mean_count, int) else -1,       # unused under force_all_rays=True (:219-222)
            True, 128, True, 0., self.max_steps, False)
        rgbs, sigmas = self.model(xyzs, dirs=dirs, **kwargs)
        if self.density_scale != 1.0:             # x * 1.0 == x exactly: skip the 16 MB pass and its backward (renderer.py:225)
            sigmas = sigmas * self.density_scale
        weights_sum, depth, image = raymarching.composite_rays_train(sigmas, rgbs, deltas, rays_info, self.t_thresh, False)
        return weights_sum, depth, image, nears, fars

    def render_train(self, rays_o, rays_d, **kwargs):
        """renderer.py:196-235"""
        weights_sum, depth, image, nears, fars = self.render_train_raw(rays_o, rays_d, **kwargs)
        classes = image[:, 3:]
        image = image[:, :3]
        image = image + (1 - weights_sum).unsqueeze(-1)
        depth = torch.clamp(depth - nears, min=0) / (fars - nears)
        return image, depth, classes

    def render_test(self, rays_o, rays_d, sync_every=4, **kwargs):
        """renderer.py:237-293.  sync_every = 1 is the reference's loop (the alive-ray count is read back every iteration,
        `rays_alive[rays_alive >= 0]`, renderer.py:284).  sync_every = k > 1 compacts on the device and refreshes the
        host-side count only every k-th iteration: in between, the list keeps its previous length with -1 in the dead
        slots (march_rays / composite_rays skip them) and n_step is derived from the stale (larger) count.  n_step only
        partitions a ray's samples over launches, so the image is the same; the loop sheds 1 - 1/k of its host syncs."""
        nears, fars = raymarching.near_far_from_aabb(rays_o, rays_d, self.aabb, self.min_near)
        N = rays_o.shape[0]
        weights_sum = torch.zeros(N, dtype=torch.float32, device=self.device)
        depth = torch.zeros(N, dtype=torch.float32, device=self.device)
        image = torch.zeros(N, self.raymarch_channels, dtype=torch.float32, device=self.device)
        n_alive = N
        rays_alive = torch.arange(n_alive, dtype=torch.int32, device=self.device)
        rays_t = nears.clone()[:, None]
        with tcnn.cache_half_params():          # the weights do not change inside one frame
            return self._render_loop(rays_o, rays_d, nears, fars, N, n_alive, rays_alive, rays_t, weights_sum, depth, image,
                                     sync_every, kwargs)

    def _render_loop(self, rays_o, rays_d, nears, fars, N, n_alive, rays_alive, rays_t, weights_sum, depth, image, sync_every,
                     kwargs):
        step, it, cnt = 0, 0, None
        while step < self.max_steps:
            if sync_every <= 1:
                n_alive = len(rays_alive)
            elif cnt is not None and it % sync_every == 0:
                n_alive = int(cnt.item())                 # the only host sync of the fast loop
                rays_alive = rays_alive[:n_alive]
            if n_alive <= 0:
                break
            n_step = max(min(N // n_alive, 8), 1)
            xyzs, dirs, deltas = raymarching.march_rays(
                n_alive, n_step, rays_alive, rays_t, rays_o, rays_d, None, self.bound, self.density_bitfield,
                self.cascade, self.grid_size, nears, fars, 128, False, 0., self.max_steps, False)
            rgbs, sigmas = self.model(xyzs, dirs=dirs, **kwargs)
            if self.density_scale != 1.0:         # renderer.py:278
                sigmas = sigmas * self.density_scale
            raymarching.composite_rays(n_alive, n_step, rays_alive, rays_t, sigmas, rgbs, deltas, False, weights_sum,
                                       depth, image, self.t_thresh)
            if sync_every <= 1:
                rays_alive = rays_alive[rays_alive >= 0]
            else:
                rays_alive, cnt = raymarching.compact_rays_alive_nosync(rays_alive)
            step += n_step
            it += 1
        classes = image[:, 3:]
        image = image[:, :3]
        image = image + (1 - weights_sum).unsqueeze(-1)
        depth = torch.clamp(depth - nears, min=0) / (fars - nears)
        return image, depth, classes

    # ------------------------------------------------------------------------------------------------------------------
    # device-driven inference loop (SURVEY.md 8f NEXT-2)
    # ------------------------------------------------------------------------------------------------------------------
    def _graph_state(self, N, row_budget):
        """Static buffers + the captured two-iteration CUDA graph for frames of N rays (built on first use)."""
        st = getattr(self, '_gs', None)
        if st is not None and st['N'] == N and st['dev'] == self.device and st['xform_ptr'] == self.model._xform.data_ptr() \
                and st['budget'] == row_budget:
            return st
        dev, Cch = self.device, self.raymarch_channels
        cap = (row_budget + 127) // 128 * 128              # rows per iteration never exceed the budget (n_alive * n_step <= budget)
        f32, i32 = torch.float32, torch.int32
        st = {'N': N, 'dev': dev, 'cap': cap, 'budget': row_budget, 'graph': None, 'xform_ptr': self.model._xform.data_ptr(),
              'rays_o': torch.empty(N, 3, dtype=f32, device=dev), 'rays_d': torch.empty(N, 3, dtype=f32, device=dev),
              'nears': torch.empty(N, dtype=f32, device=dev), 'fars': torch.empty(N, dtype=f32, device=dev),
              'rays_t': torch.empty(N, 1, dtype=f32, device=dev),
              'alive': [torch.empty(N, dtype=i32, device=dev), torch.empty(N, dtype=i32, device=dev)],
              'ctl': torch.zeros(8, dtype=i32, device=dev),
              'xyzs': torch.zeros(cap, 3, dtype=f32, device=dev), 'dirs': torch.zeros(cap, 3, dtype=f32, device=dev),
              'deltas': torch.zeros(cap, 4, dtype=f32, device=dev),
              'enc_d': torch.zeros(cap, 32, dtype=torch.float16, device=dev), 'enc_c': torch.zeros(cap, 32, dtype=torch.float16, device=dev),
              'c1': torch.zeros(cap, 16, dtype=torch.float16, device=dev),
              'sigmas': torch.zeros(cap, 1, dtype=f32, device=dev), 'rgbs': torch.zeros(cap, Cch, dtype=f32, device=dev),
              'weights_sum': torch.empty(N, dtype=f32, device=dev), 'depth': torch.empty(N, dtype=f32, device=dev),
              'image': torch.empty(N, Cch, dtype=f32, device=dev),
              # both fp16 tables of this frame, interleaved [row][table][2]: one 8-byte gather per corner serves both
              'pair': torch.empty(self.model.x_density_embedder.embeddings.shape[0], 2, 2, dtype=torch.float16, device=dev),
              'w': {n: torch.empty(getattr(self.model, n).params.numel(), dtype=torch.float16, device=dev)
                    for n in ('density_net', 'class_net', 'color1_net', 'color2_net')},
              'scratch': torch.empty(int(raymarching.L.lib().nrf_march_scratch_bytes(N)) + 64, dtype=torch.uint8, device=dev)}
        self._gs = st
        return st

    def _graph_iteration(self, st, a_in, a_out):
        """One iteration of renderer.py:249-286 with every count on the device (ctl) and every launch sized for the cap."""
        L = raymarching.L
        lib, m = L.lib(), self.model
        s = torch.cuda.current_stream(self.device).cuda_stream
        N, cap, ctl = st['N'], st['cap'], st['ctl']
        enc = m.x_density_embedder
        S = float(np.float32(np.log2(enc.per_level_scale)))
        rows = ctl.data_ptr() + 8                                  # &ctl[2] = n_alive * n_step
        L.check(lib.nrf_march_rays_dev(ctl.data_ptr(), N, a_in.data_ptr(), st['rays_t'].data_ptr(), st['rays_o'].data_ptr(),
                                       st['rays_d'].data_ptr(), float(self.bound), 0.0, int(self.max_steps), int(self.cascade),
                                       int(self.grid_size), self.density_bitfield.data_ptr(), st['fars'].data_ptr(),
                                       st['xyzs'].data_ptr(), None, st['deltas'].data_ptr(), s), 'march_rays_dev')      # no direction input
        L.check(lib.nrf_grid_encode_forward_pair(st['xyzs'].data_ptr(), st['pair'].data_ptr(), enc.offsets.data_ptr(),
                                                 st['enc_d'].data_ptr(), st['enc_c'].data_ptr(), cap, enc.num_levels, S,
                                                 int(enc.base_resolution), enc.gridtype_id, int(enc.align_corners), 0,
                                                 L.DTYPE_F16, m._xform.data_ptr(), rows, st['deltas'].data_ptr(), s),
                'grid_encode_forward_pair')

        def mlp(net, x, y, col, n_out, act):
            L.check(lib.nrf_mlp_forward_dev(x.data_ptr(), L.dtype_code(x.dtype), st['w'][net].data_ptr(), cap,
                                            getattr(m, net).n_input_dims, n_out, getattr(m, net).n_hidden_layers, 64,
                                            getattr(m, net).hidden_act, act, y.data_ptr() + col * y.element_size(),
                                            L.dtype_code(y.dtype), y.shape[1], rows, s), 'mlp_forward_dev')
        if tcnn._field_fusable(st['enc_d'], st['enc_c'], m.density_net, m.class_net, m.color1_net, m.color2_net) and lib.nrf_mlp_get_mode() == 0:
            L.check(lib.nrf_field_forward(st['enc_d'].data_ptr(), st['enc_c'].data_ptr(), st['w']['density_net'].data_ptr(),
                                          st['w']['class_net'].data_ptr(), st['w']['color1_net'].data_ptr(), st['w']['color2_net'].data_ptr(),
                                          cap, m.class_dim, st['sigmas'].data_ptr(), st['rgbs'].data_ptr(), st['rgbs'].shape[1], None, rows, s),
                    'field_forward')
        else:
            mlp('density_net', st['enc_d'], st['sigmas'], 0, 1, L.ACT['trunc_exp'])
            mlp('color1_net', st['enc_c'], st['c1'], 0, 16, m.color1_net.out_act)
            mlp('color2_net', st['c1'], st['rgbs'], 0, 3, m.color2_net.out_act)
            mlp('class_net', st['enc_c'], st['rgbs'], 3, m.class_dim, m.class_net.out_act)
        if self.density_scale != 1.0:
            st['sigmas'].mul_(self.density_scale)
        L.check(lib.nrf_composite_rays_dev(ctl.data_ptr(), N, float(self.t_thresh), a_in.data_ptr(), st['rays_t'].data_ptr(),
                                           st['sigmas'].data_ptr(), st['rgbs'].data_ptr(), st['deltas'].data_ptr(),
                                           self.raymarch_channels, st['weights_sum'].data_ptr(), st['depth'].data_ptr(),
                                           st['image'].data_ptr(), s), 'composite_rays_dev')
        L.check(lib.nrf_compact_alive_dev(ctl.data_ptr(), N, a_in.data_ptr(), a_out.data_ptr(), st['scratch'].data_ptr(), s),
                'compact_alive_dev')

    @torch.no_grad()
    def render_test_graph(self, rays_o, rays_d, check_every=2, steps_per_iteration=8):
        """render_test with the loop driven from the device: the per-iteration host logic of renderer.py:249-286 (alive
        count, n_step, buffer sizes) lives in a control block updated by the compaction kernel, every launch is sized for
        the cap, and a PAIR of iterations (the alive list ping-pongs between two buffers) is captured once in a CUDA graph
        and replayed; the host only reads the alive count every `check_every` replays (a replay = two iterations; with 8
        steps per iteration a frame is ~4 replays, so 2 keeps the replays wasted after the last ray died to at most one).  Same kernels and numerics as
        render_test; needs the fused-head model (default) under AMP-style fp16 tables.

        steps_per_iteration: the reference marches n_step = clamp(N // n_alive, 1, 8) samples per ray and iteration, i.e. ONE
        while most rays are alive, so a frame is ~56 iterations that each re-read and re-write every ray's accumulators
        (renderer.py:253).  n_step only partitions a ray's samples over iterations -- composite_rays stops at the terminating
        sample inside the block -- so the loop here budgets steps_per_iteration * N rows per iteration instead
        (n_step = clamp(budget // n_alive, 1, 8)): 8x fewer iterations, accumulator round trips and compaction passes, for
        at most n_step - 1 wasted samples per ray per frame (default 8: 12.4 ms per 1008x756 frame; 4: 13.0; 1: 17).
        steps_per_iteration=1 is the reference's schedule."""
        m = self.model
        if not (getattr(m, 'fused_heads', False) and m.class_dim + 3 == self.raymarch_channels):
            raise RuntimeError('render_test_graph needs the fused-head StyleTCNerf')
        if m._dual is None:
            m._dual = bool(same_geometry(m.x_density_embedder, m.x_color_embedder))
            m._xform = torch.cat([m.bbox_min, m.bbox_size, m.bbox_min.new_ones(1)]).contiguous()
        if not m._dual:
            raise RuntimeError('render_test_graph needs the two encoders to share their geometry')
        rays_o = rays_o.float().contiguous().view(-1, 3)
        rays_d = rays_d.float().contiguous().view(-1, 3)
        N = rays_o.shape[0]
        budget = N * max(1, min(int(steps_per_iteration), 8))
        st = self._graph_state(N, budget)
        st['rays_o'].copy_(rays_o); st['rays_d'].copy_(rays_d)
        nears, fars = raymarching.near_far_from_aabb(st['rays_o'], st['rays_d'], self.aabb, self.min_near)
        st['nears'].copy_(nears); st['fars'].copy_(fars)
        st['rays_t'].copy_(nears[:, None])
        st['weights_sum'].zero_(); st['depth'].zero_(); st['image'].zero_()
        st['alive'][0].copy_(torch.arange(N, dtype=torch.int32, device=self.device))
        n0 = max(min(budget // N, 8), 1)
        st['ctl'].copy_(torch.tensor([N, n0, N * n0, 0, budget, self.max_steps, 0, 1], dtype=torch.int32))      # ctl[7] = 1: step-major sample rows
        for i, e in enumerate((m.x_density_embedder, m.x_color_embedder)):       # fp16 tables / weights for this frame
            shadow = current_half_copy(e.embeddings)
            st['pair'][:, i].copy_(shadow if shadow is not None else e.embeddings.detach())
        for n in st['w']:
            st['w'][n].copy_(getattr(m, n).params.detach())
        if st['graph'] is None:
            # the first pair runs eagerly (it is real work: the two largest iterations, and it initialises the kernels'
            # attributes), then the same pair is captured for every later replay
            self._graph_iteration(st, st['alive'][0], st['alive'][1])
            self._graph_iteration(st, st['alive'][1], st['alive'][0])
            torch.cuda.synchronize(self.device)
            g = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g):
                self._graph_iteration(st, st['alive'][0], st['alive'][1])
                self._graph_iteration(st, st['alive'][1], st['alive'][0])
            st['graph'] = g
        it = 0
        while True:
            if it % check_every == 0 and int(st['ctl'][0].item()) <= 0:      # the loop's only host sync
                break
            st['graph'].replay()
            it += 1
        image = st['image']
        classes = image[:, 3:].clone()
        rgb = image[:, :3] + (1 - st['weights_sum']).unsqueeze(-1)
        depth = torch.clamp(st['depth'] - st['nears'], min=0) / (st['fars'] - st['nears'])
        return rgb, depth, classes
