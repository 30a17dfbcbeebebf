"""Fused optimizer for the field's parameters: GradScaler + Adam(eps=1e-15) + LambdaLR + torch_ema in one pass per
tensor on the device (csrc/optim.cu), mirroring trainers/base.py:216-229,420-426 without any host synchronisation.

It also maintains the fp16 copy of each hash table, which `GridEncoder` picks up under autocast instead of re-casting
the 48 MB fp32 table on every forward (gridencoder/grid.py:42-43 does `embeddings.to(torch.half)` per call).
"""
import struct

import torch

from . import _lib as L


class FusedAdamEMA:
    def __init__(self, params, lr=0.01, betas=(0.9, 0.999), eps=1e-15, lr_decay_steps=30000, ema_decay=0.95,
                 init_scale=65536.0, growth_factor=2.0, backoff_factor=0.5, growth_interval=2000, enable_amp=True,
                 half_copy_min_numel=1 << 20):
        self.params = [p for p in params]
        dev = self.params[0].device
        self.lr, self.betas, self.eps, self.lr_decay_steps = lr, betas, eps, float(lr_decay_steps)
        self.ema_decay = ema_decay
        self.growth_factor, self.backoff_factor, self.growth_interval = growth_factor, backoff_factor, growth_interval
        self.enable_amp = enable_amp
        self.exp_avg = [torch.zeros_like(p) for p in self.params]
        self.exp_avg_sq = [torch.zeros_like(p) for p in self.params]
        self.ema = [p.detach().clone() for p in self.params] if ema_decay is not None else None
        self.num_updates = 0
        nbytes = int(L.lib().nrf_opt_state_bytes())
        raw = struct.pack('fiii', float(init_scale if enable_amp else 1.0), 0, 0, 0) + b'\0' * (nbytes - 16)
        self.state = torch.frombuffer(bytearray(raw), dtype=torch.uint8).to(dev)
        # views of the state block (device tensors; reading them on the host synchronises -- only for logging)
        self.scale = self.state[:4].view(torch.float32)
        self.good_steps = self.state[12:16].view(torch.int32)
        # fp16 shadow copies of the big tables, registered on the parameter for GridEncoder to find
        self.half = []
        for p in self.params:
            # big tensors (hash tables): fp16 copy only under AMP; small ones (the MLPs' flat params): always -- the
            # tensor-core kernels consume fp16 weights in every mode (nerfstyle_b200.tcnn.half_params)
            if (enable_amp and p.numel() >= half_copy_min_numel) or p.numel() < half_copy_min_numel:
                h = p.detach().to(torch.float16)
                p._nrf_half_copy = h
                self.half.append(h)
            else:
                self.half.append(None)

    @torch.no_grad()
    def refresh_half_copies(self):
        """Call after writing the parameters by any other means (loading a checkpoint, swapping in the EMA weights)."""
        for p, h in zip(self.params, self.half):
            if h is not None:
                h.copy_(p.detach())

    def scale_loss(self, loss):
        return loss * self.scale if self.enable_amp else loss

    def zero_grad(self):
        for p in self.params:
            p.grad = None

    @torch.no_grad()
    def step(self):
        lib = L.lib()
        dev = self.params[0].device
        self.num_updates += 1
        if self.ema is not None:
            decay = min(self.ema_decay, (1 + self.num_updates) / (10 + self.num_updates))    # torch_ema, use_num_updates
            omd = 1.0 - decay
        else:
            omd = 0.0
        with torch.cuda.device(dev):
            st = torch.cuda.current_stream(dev).cuda_stream
            grads = []
            for p in self.params:
                g = p.grad
                if g is None:
                    grads.append(None)
                    continue
                if g.dtype != torch.float32 or not g.is_contiguous():
                    g = g.float().contiguous()
                grads.append(g)
                L.check(lib.nrf_grads_check(g.data_ptr(), g.numel(), self.state.data_ptr(), st), 'grads_check')
            for i, p in enumerate(self.params):
                if grads[i] is None:
                    continue
                L.check(lib.nrf_adam_step(p.data_ptr(), grads[i].data_ptr(), self.exp_avg[i].data_ptr(),
                                          self.exp_avg_sq[i].data_ptr(), L.ptr(self.ema[i]) if self.ema is not None else None,
                                          L.ptr(self.half[i]), p.numel(), self.state.data_ptr(), self.lr, self.lr_decay_steps,
                                          self.betas[0], self.betas[1], self.eps, omd, st), 'adam_step')
            L.check(lib.nrf_scaler_update(self.state.data_ptr(), self.growth_factor, self.backoff_factor,
                                          int(self.growth_interval) if self.enable_amp else (1 << 30), st), 'scaler_update')
