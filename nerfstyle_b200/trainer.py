"""One reconstruction training step, mirroring Trainer.run_iter / calc_loss of the reference
(trainers/base.py:251-304, 396-426): autocast(fp16) render of a ray batch -> MSE + lambda * CE(class) ->
GradScaler-scaled backward -> Adam(eps=1e-15, betas (0.9,0.999)) -> LambdaLR decay -> EMA of the parameters.

Data-parallel form: each rank renders its shard of the batch with the loss scaled by
n_local / n_global, then the hash-table and MLP gradients are summed with one NCCL all-reduce per tensor
group before the (identical) optimizer step on every rank.
"""
import torch
import torch.nn.functional as F
from torch.autograd import Function


class _ReconLoss(Function):
    """White background (renderer.py:229-232) + MSE + class_lambda * cross-entropy (trainers/base.py:251-304) and their
    gradients w.r.t. the compositing outputs in ONE launch (csrc/loss.cu) instead of ~25 elementwise / reduction / slicing
    kernels and their backward twins.  Returns a [3] tensor: (total, mse, ce)."""

    _scratch = {}

    @staticmethod
    def forward(ctx, weights_sum, image, target_rgb, target_cls, class_lambda):
        from . import _lib as L
        N, C = image.shape
        image, weights_sum = image.float().contiguous(), weights_sum.float().contiguous()
        target_rgb = target_rgb.float().contiguous()
        target_cls = target_cls.to(torch.int64).contiguous() if C > 3 else None
        out = torch.empty(3, dtype=torch.float32, device=image.device)
        g_img = torch.empty_like(image)
        g_ws = torch.empty_like(weights_sum)
        lib = L.lib()
        key = (image.device, int(lib.nrf_recon_loss_scratch_bytes(N)))
        scratch = _ReconLoss._scratch.get(key)
        if scratch is None:          # zero once: every call leaves the arrival counter at zero again
            scratch = _ReconLoss._scratch[key] = torch.zeros(key[1] // 8 + 1, dtype=torch.float64, device=image.device)
        with torch.cuda.device(image.device):
            L.check(lib.nrf_recon_loss(image.data_ptr(), weights_sum.data_ptr(), target_rgb.data_ptr(), L.ptr(target_cls), N, C,
                                       float(class_lambda), out.data_ptr(), g_img.data_ptr(), g_ws.data_ptr(),
                                       scratch.data_ptr(), L.stream_of(image)), 'recon_loss')
        ctx.save_for_backward(g_ws, g_img)
        return out

    @staticmethod
    def backward(ctx, g):
        g_ws, g_img = ctx.saved_tensors
        s = g[0]                                   # only the total (out[0]) carries the training gradient
        return g_ws * s, g_img * s, None, None, None


def recon_loss(weights_sum, image, target_rgb, target_cls, class_lambda):
    """(total, mse, ce) as 0-dim tensors; differentiable through `total` w.r.t. weights_sum and image."""
    out = _ReconLoss.apply(weights_sum, image, target_rgb, target_cls, class_lambda)
    return out[0], out[1].detach(), out[2].detach()


class TrainStep:
    def __init__(self, renderer, lr=0.01, mlp_lr=None, lr_decay=30000, ema_decay=0.95, class_lambda=0.001, enable_amp=True,
                 fused_adam=True, world_size=1, fused_optimizer=True, rank=None, shard_optimizer=True, pair_tables=True,
                 fused_loss=True):
        self.renderer = renderer
        self.model = renderer.model
        params = list(self.model.parameters())
        self.params = params
        self.fused_loss = fused_loss
        self.enable_amp = enable_amp
        self.class_lambda = class_lambda
        self.world_size = world_size
        self.iter_ctr = 0
        self._side = None
        self.loss_ready = None
        self.fused = None
        if fused_optimizer:
            # GradScaler + Adam + LambdaLR + EMA + fp16 table copies in one device pass per tensor, no host sync
            from .optim import FusedAdamEMA
            if rank is None:
                import torch.distributed as dist
                rank = dist.get_rank() if (world_size > 1 and dist.is_initialized()) else 0
            # with world_size > 1 the optimizer owns the gradient exchange (reduce-scatter / sharded Adam / all-gather)
            # mlp_lr: the reference's optional second parameter group (trainers/base.py:199-212, keywords2 -> lr 0.005)
            lrs = [mlp_lr if (mlp_lr is not None and 'net' in n) else lr for n, _ in self.model.named_parameters()]
            self.fused = FusedAdamEMA(params, lr=lrs, eps=1e-15, lr_decay_steps=lr_decay, ema_decay=ema_decay,
                                      enable_amp=enable_amp, world_size=world_size, rank=rank, shard_big=shard_optimizer,
                                      pair_tables=pair_tables)
            self.ema = self.fused.ema
            return
        from .optim import optimizer_of
        for p in params:                    # a fused optimizer from an earlier stage lets go of the parameters first
            prev = optimizer_of(p)
            if prev is not None:
                prev.detach()
        groups = [{'params': params}]
        if mlp_lr is not None:
            named = list(self.model.named_parameters())
            groups = [{'params': [p for n, p in named if 'net' not in n]}, {'params': [p for n, p in named if 'net' in n], 'lr': mlp_lr}]
        self.optim = torch.optim.Adam(groups, lr=lr, betas=(0.9, 0.999), eps=1e-15, fused=fused_adam)
        self.scheduler = torch.optim.lr_scheduler.LambdaLR(
            self.optim, (lambda it: 0.1 ** (it / lr_decay)) if lr_decay > 0 else (lambda it: 1.0))
        self.scaler = torch.amp.GradScaler('cuda', enabled=enable_amp)
        self.enable_amp = enable_amp
        self.class_lambda = class_lambda
        self.ema_decay = ema_decay
        self.ema = [p.detach().clone() for p in params] if ema_decay > 0 else None
        self.world_size = world_size
        self.iter_ctr = 0

    @staticmethod
    def reserve_workspace(device, nbytes=12 << 30):
        """One arena for the step's sample buffers: allocate-and-release `nbytes` once so that the caching allocator holds
        a single large block it can split for every later request.  The number of samples per step drifts as the field
        trains, and without the arena each new high-water mark costs a cudaMalloc (a device-wide synchronisation) in the
        middle of training; a 180 GB part has the room."""
        torch.empty(int(nbytes), dtype=torch.uint8, device=device)

    def loss_fn(self, image, classes, target_rgb, target_cls):
        mse = torch.mean((image - target_rgb) ** 2)
        cls = F.cross_entropy(classes, target_cls) * self.class_lambda
        return mse + cls, mse

    def allreduce_grads(self):
        from .parallel import allreduce_grads
        return allreduce_grads(self.params, self.world_size)

    def step(self, rays_o, rays_d, target_rgb, target_cls, n_global=None, loss_host=None):
        """rays_* [n,3], target_rgb [n,3] f32, target_cls [n] int64 -- all on the device.  Returns the loss tensor
        (device; no host sync here beyond the one inside march_rays_train).

        loss_host: optional pinned f32[1] host tensor.  The loss is copied into it on a side stream as soon as the
        FORWARD pass has produced it (the backward + optimizer kernels are still being enqueued / executed), and the
        event to wait on before reading it is returned as `self.loss_ready`; a logging read-back therefore never
        stalls the step behind its own backward."""
        n_local = rays_o.shape[0]
        n_global = n_global or n_local * self.world_size
        with torch.autocast('cuda', dtype=torch.float16, enabled=self.enable_amp):
            if self.fused_loss and rays_o.is_cuda:
                weights_sum, _, image_raw, _, _ = self.renderer.render_train_raw(rays_o, rays_d)
                loss, mse, _ = recon_loss(weights_sum, image_raw, target_rgb, target_cls, self.class_lambda)
            else:
                image, depth, classes = self.renderer.render_train(rays_o, rays_d)
                loss, mse = self.loss_fn(image, classes, target_rgb, target_cls)
        if loss_host is not None:
            if self._side is None:
                self._side = torch.cuda.Stream(device=rays_o.device)
            produced = torch.cuda.Event()
            produced.record()
            ld = loss.detach()
            with torch.cuda.stream(self._side):
                self._side.wait_event(produced)
                loss_host.copy_(ld.reshape(1), non_blocking=True)
                self.loss_ready = torch.cuda.Event()
                self.loss_ready.record()
            ld.record_stream(self._side)
        back = loss * (n_local / n_global) if self.world_size > 1 else loss
        if self.fused is not None:
            self.fused.zero_grad()
            self.fused.scale_loss(back).backward()
            self.fused.step()            # includes the gradient exchange when world_size > 1
            self.iter_ctr += 1
            return loss.detach()
        self.optim.zero_grad(set_to_none=True)
        self.scaler.scale(back).backward()
        self.allreduce_grads()
        self.scaler.step(self.optim)
        old_scale = self.scaler.get_scale() if self.enable_amp else 1.0
        self.scaler.update()
        if not self.enable_amp or old_scale <= self.scaler.get_scale():
            self.scheduler.step()
        if self.ema is not None:
            with torch.no_grad():
                n = self.iter_ctr + 1
                decay = min(self.ema_decay, (1 + n) / (10 + n))          # torch_ema (use_num_updates=True), as the fused path
                torch._foreach_mul_(self.ema, decay)
                torch._foreach_add_(self.ema, [p.detach() for p in self.params], alpha=1.0 - decay)
        self.iter_ctr += 1
        return loss.detach()
