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
                 half_copy_min_numel=1 << 20, world_size=1, rank=0, shard_big=True, pair_tables=True):
        """world_size > 1 makes step() own the gradient exchange of the data-parallel step:
        * small tensors (the MLPs): one flattened all-reduce, replicated update;
        * big tensors (the hash tables), under AMP: REDUCE-SCATTER of the gradient, Adam / EMA on this rank's 1/world
          shard only (optimizer state is allocated for the shard), then ALL-GATHER of the fp16 table copy -- the only
          form the next forward reads.  That moves 3/4 of an all-reduce's bytes and divides the optimizer pass by world.
          The fp32 master table is then current only inside each rank's shard (gather_master() rebuilds it, e.g. for a
          checkpoint); without AMP the tables are read in fp32 and the big tensors fall back to all-reduce.

        pair_tables (under AMP): when two big tensors share one [T, 2] shape (the model's two hash tables),
        their fp16 copies live in ONE interleaved buffer [T][table][2] and their gradients are accumulated by the
        paired scatter kernel into ONE interleaved f32 buffer (`grad_pair`) that this optimizer reads directly -- the
        tables' `.grad` stays None (use grad_of(p) to look at a gradient).  A corner of both tables is then one
        vector gather / one 16-byte reduction in the hash-grid kernels (nrf_grid_encode_forward_pair / _backward_pair)."""
        self.params = [p for p in params]
        dev = self.params[0].device
        self.world_size, self.rank = int(world_size), int(rank)
        self.lr, self.betas, self.eps, self.lr_decay_steps = lr, betas, eps, float(lr_decay_steps)
        self.ema_decay = ema_decay
        self.growth_factor, self.backoff_factor, self.growth_interval = growth_factor, backoff_factor, growth_interval
        self.enable_amp = enable_amp
        # shard plan: (lo, hi) element range of this rank for every sharded tensor, None for replicated ones
        self.shard = []
        for p in self.params:
            n = p.numel()
            if self.world_size > 1 and shard_big and enable_amp and n >= half_copy_min_numel and n % self.world_size == 0:
                per = n // self.world_size
                self.shard.append((self.rank * per, (self.rank + 1) * per))
            else:
                self.shard.append(None)

        def state_like(p, sh, init=None):
            if sh is None:
                return torch.zeros_like(p) if init is None else p.detach().clone()
            flat = p.detach().reshape(-1)[sh[0]:sh[1]]
            return torch.zeros_like(flat) if init is None else flat.clone()
        # the pair of same-shaped tables sharing interleaved buffers (indices into self.params), or None
        self.pair_idx = None
        if pair_tables and enable_amp:
            groups = {}
            for i, p in enumerate(self.params):
                if p.numel() >= half_copy_min_numel and p.dim() == 2 and p.shape[1] == 2:
                    groups.setdefault(tuple(p.shape), []).append(i)
            twins = [g for g in groups.values() if len(g) == 2]
            if len(twins) == 1:
                big = twins[0]
                if all(self.shard[i] is None or self.shard[i][0] % 2 == 0 for i in big) and self.shard[big[0]] == self.shard[big[1]]:
                    self.pair_idx = tuple(big)
        self.grad_pair = None           # [T, 2, 2] f32, allocated by the first paired backward
        self.grad_pair_valid = False    # holds this step's gradients (zero_grad() invalidates; the next backward clears it)
        self.grad_shard_pair = None
        self.exp_avg = [state_like(p, sh) for p, sh in zip(self.params, self.shard)]
        self.exp_avg_sq = [state_like(p, sh) for p, sh in zip(self.params, self.shard)]
        self.ema = [state_like(p, sh, 'copy') for p, sh in zip(self.params, self.shard)] if ema_decay is not None else None
        self.grad_shard = [torch.empty(sh[1] - sh[0], dtype=torch.float32, device=dev) if sh is not None else None
                           for sh in self.shard]
        self.num_updates = 0
        nbytes = int(L.lib().nrf_opt_state_bytes())
        raw = struct.pack('fiii', float(init_scale if enable_amp else 1.0), 0, 0, 0) + b'\0' * (nbytes - 16)
        self.state = torch.frombuffer(bytearray(raw), dtype=torch.uint8).to(dev)
        # views of the state block (device tensors; reading them on the host synchronises -- only for logging)
        self.scale = self.state[:4].view(torch.float32)
        self.good_steps = self.state[12:16].view(torch.int32)
        # fp16 shadow copies of the big tables, registered on the parameter for GridEncoder to find
        self.half = []
        self.half_pair = None
        if self.pair_idx is not None:
            a, b = (self.params[i] for i in self.pair_idx)
            self.half_pair = torch.stack([a.detach().to(torch.float16), b.detach().to(torch.float16)], dim=1).contiguous()
        for i, p in enumerate(self.params):
            if self.pair_idx is not None and i in self.pair_idx:
                e = self.pair_idx.index(i)
                h = self.half_pair[:, e]                 # strided view: the interleaved buffer is the only fp16 copy
                p._nrf_half_copy = h
                p._nrf_half_pair = (self.half_pair, e)
                p._nrf_grad_sink = (self, e)
                self.half.append(h)
                continue
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
        self.grad_pair_valid = False

    def grad_pair_buffer(self):
        """The interleaved gradient buffer the paired scatter accumulates into (called by the dual encoder's backward)."""
        a = self.params[self.pair_idx[0]]
        if self.grad_pair is None:
            self.grad_pair = torch.zeros(a.shape[0], 2, 2, dtype=torch.float32, device=a.device)
        elif not self.grad_pair_valid:
            self.grad_pair.zero_()
        self.grad_pair_valid = True
        return self.grad_pair

    def grad_of(self, p):
        """The (scaled) gradient of a parameter wherever it lives: p.grad, or its slot of the interleaved pair buffer."""
        if p.grad is not None:
            return p.grad
        sink = getattr(p, '_nrf_grad_sink', None)
        if sink is not None and sink[0] is self and self.grad_pair_valid:
            return self.grad_pair[:, sink[1]]
        return None

    @torch.no_grad()
    def gather_master(self):
        """Rebuild the full fp32 master copy of every sharded tensor on every rank (checkpointing)."""
        from . import parallel
        for p, sh in zip(self.params, self.shard):
            if sh is not None:
                flat = p.data.view(-1)
                parallel.all_gather_shards(flat, flat[sh[0]:sh[1]].clone(), self.world_size)

    @torch.no_grad()
    def step(self):
        from . import parallel
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
                if g is not None and (g.dtype != torch.float32 or not g.is_contiguous()):
                    g = g.float().contiguous()
                grads.append(g)
            # gradients of the paired tables arrive in the interleaved buffer unless someone produced .grad the usual way
            pair_grads = self.pair_idx is not None and self.grad_pair_valid and all(grads[i] is None for i in self.pair_idx)
            pair_src = None                    # (tensor holding this rank's interleaved gradient rows)
            if pair_grads:
                pair_src = self.grad_pair
            if self.world_size > 1:
                # ---- gradient exchange: reduce-scatter the sharded tensors, one flattened all-reduce for the rest
                for i, sh in enumerate(self.shard):
                    if sh is not None and grads[i] is not None:
                        parallel.reduce_scatter_sum(grads[i].view(-1), self.grad_shard[i], self.world_size, self.rank)
                        grads[i] = self.grad_shard[i]
                if pair_grads:
                    if self.shard[self.pair_idx[0]] is not None:
                        # row shards of the two tables coincide: ONE reduce-scatter of the interleaved buffer
                        flat = self.grad_pair.view(-1)
                        if self.grad_shard_pair is None:
                            self.grad_shard_pair = torch.empty(flat.numel() // self.world_size, dtype=torch.float32, device=dev)
                        parallel.reduce_scatter_sum(flat, self.grad_shard_pair, self.world_size, self.rank)
                        pair_src = self.grad_shard_pair
                    else:
                        parallel.allreduce_tensors([self.grad_pair], self.world_size)
                parallel.allreduce_tensors([g for g, sh in zip(grads, self.shard) if sh is None and g is not None], self.world_size)
            for g in grads:
                if g is not None:
                    L.check(lib.nrf_grads_check(g.data_ptr(), g.numel(), self.state.data_ptr(), st), 'grads_check')
            if pair_grads:
                L.check(lib.nrf_grads_check(pair_src.data_ptr(), pair_src.numel(), self.state.data_ptr(), st), 'grads_check')
            if self.world_size > 1 and any(sh is not None for sh in self.shard):
                # an inf seen in ANY rank's shard skips the step everywhere (the replicated GradScaler decision)
                parallel.allreduce_max_int(self.state[4:8].view(torch.int32), self.world_size)
            gathered_pair = False
            if pair_grads:
                # both tables in one pass over the interleaved gradient rows (16-byte loads) and fp16 rows (8-byte stores)
                ia, ib = self.pair_idx
                sh = self.shard[ia]
                lo, n = (0, self.params[ia].numel()) if sh is None else (sh[0], sh[1] - sh[0])
                L.check(lib.nrf_adam_step_pair(
                    self.params[ia].data_ptr() + 4 * lo, self.params[ib].data_ptr() + 4 * lo, pair_src.data_ptr(),
                    self.exp_avg[ia].data_ptr(), self.exp_avg[ib].data_ptr(), self.exp_avg_sq[ia].data_ptr(),
                    self.exp_avg_sq[ib].data_ptr(), L.ptr(self.ema[ia]) if self.ema is not None else None,
                    L.ptr(self.ema[ib]) if self.ema is not None else None, self.half_pair.data_ptr() + (lo // 2) * 8, n // 2,
                    self.state.data_ptr(), self.lr, self.lr_decay_steps, self.betas[0], self.betas[1], self.eps, omd, st),
                    'adam_step_pair')
            for i, p in enumerate(self.params):
                paired = self.pair_idx is not None and i in self.pair_idx
                if grads[i] is None or (paired and pair_grads):
                    continue
                sh = self.shard[i]
                lo, n = (0, p.numel()) if sh is None else (sh[0], sh[1] - sh[0])
                p_ptr = p.data_ptr() + 4 * lo
                g_ptr, g_stride, h_stride = grads[i].data_ptr(), 1, 1
                if paired:         # a paired table whose gradient arrived as an ordinary .grad: only the fp16 copy is strided
                    h_ptr, h_stride = self.half_pair.data_ptr() + (lo // 2) * 8 + self.pair_idx.index(i) * 4, 2
                else:
                    h_ptr = (self.half[i].data_ptr() + 2 * lo) if self.half[i] is not None else None
                L.check(lib.nrf_adam_step_ex(p_ptr, g_ptr, self.exp_avg[i].data_ptr(), self.exp_avg_sq[i].data_ptr(),
                                             L.ptr(self.ema[i]) if self.ema is not None else None, h_ptr, n,
                                             self.state.data_ptr(), self.lr, self.lr_decay_steps, self.betas[0], self.betas[1],
                                             self.eps, omd, g_stride, h_stride, st), 'adam_step')
            if self.world_size > 1:
                # ---- the next forward reads only the fp16 table copies: gather their shards (in place)
                for i, sh in enumerate(self.shard):
                    if sh is None or (grads[i] is None and not (pair_grads and i in self.pair_idx)):
                        continue
                    if self.pair_idx is not None and i in self.pair_idx:
                        if not gathered_pair:          # param elements [lo, hi) <-> interleaved halfs [2 lo, 2 hi)
                            hflat = self.half_pair.view(-1)
                            parallel.all_gather_shards(hflat, hflat[2 * sh[0]:2 * sh[1]], self.world_size)
                            gathered_pair = True
                        continue
                    hflat = self.half[i].view(-1)
                    parallel.all_gather_shards(hflat, hflat[sh[0]:sh[1]], self.world_size)
            L.check(lib.nrf_scaler_update(self.state.data_ptr(), self.growth_factor, self.backoff_factor,
                                          int(self.growth_interval) if self.enable_amp else (1 << 30), st), 'scaler_update')
