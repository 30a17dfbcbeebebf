"""Fused optimizer for the field's parameters: GradScaler + Adam(eps=1e-15) + LambdaLR + torch_ema in one pass per
tensor on the device (csrc/optim.cu), mirroring trainers/base.py:216-229,420-426 without any host synchronisation.

It also maintains the fp16 copy of each hash table, which `GridEncoder` picks up under autocast instead of re-casting
the 48 MB fp32 table on every forward (gridencoder/grid.py:42-43 does `embeddings.to(torch.half)` per call).

Ownership protocol.  The optimizer hangs four attributes on each parameter it owns -- `_nrf_optimizer` (weak reference
to the owner), `_nrf_half_copy` (+ `_nrf_half_version`: the parameter's version counter when the copy was last made),
`_nrf_half_pair`, `_nrf_grad_sink` -- and removes them again in `detach()`; constructing a new FusedAdamEMA over a
parameter detaches the previous owner.  Readers go through `current_half_copy()`, which re-casts a copy whose
parameter has been written since (load_state_dict, `param.copy_`, anything that bumps the version counter); writes
through `param.data` are invisible to PyTorch itself and need an explicit `refresh_half_copies()` -- `ema_scope()`
does that for the one such write this module performs.
"""
import contextlib
import struct
import weakref

import torch

from . import _lib as L

_ATTRS = ('_nrf_optimizer', '_nrf_half_copy', '_nrf_half_version', '_nrf_half_pair', '_nrf_grad_sink')


def optimizer_of(p):
    """The live FusedAdamEMA that owns parameter `p`, or None."""
    ref = getattr(p, '_nrf_optimizer', None)
    opt = ref() if ref is not None else None
    return opt if (opt is not None and opt.alive) else None


def current_half_copy(p):
    """The fp16 shadow copy of `p` kept by its optimizer (None when there is none), re-cast first if `p` has been
    written since the copy was made."""
    h = getattr(p, '_nrf_half_copy', None)
    if h is None:
        return None
    opt = optimizer_of(p)
    if opt is None:                       # orphaned attributes (the owner was garbage-collected without detach())
        for a in _ATTRS:
            if hasattr(p, a):
                delattr(p, a)
        return None
    if p._version != p._nrf_half_version:
        opt.refresh_half_copies(only=p, external_write=True)
    opt.wait_pending_gather()             # the all-gather of the previous step may still be running on the side stream
    return h


def current_half_pair(p):
    """(interleaved fp16 buffer, slot) of a paired table, fresh; or None."""
    if current_half_copy(p) is None:
        return None
    return getattr(p, '_nrf_half_pair', None)


def live_grad_sink(p):
    s = getattr(p, '_nrf_grad_sink', None)
    if s is None or not s[0].alive:
        return None
    return s


def sync_for_checkpoint(module):
    """Before reading parameters for a checkpoint: a sharded optimizer keeps the fp32 master tables current only inside
    each rank's shard, so gather them (a COLLECTIVE -- every rank must take the checkpoint path)."""
    seen = set()
    for p in module.parameters():
        opt = optimizer_of(p)
        if opt is not None and id(opt) not in seen:
            seen.add(id(opt))
            if not opt.master_complete:
                opt.gather_master()


class FusedAdamEMA:
    def __init__(self, params, lr=0.01, betas=(0.9, 0.999), eps=1e-15, lr_decay_steps=30000, ema_decay=0.95,
                 init_scale=65536.0, growth_factor=2.0, backoff_factor=0.5, growth_interval=2000, enable_amp=True,
                 half_copy_min_numel=1 << 20, world_size=1, rank=0, shard_big=True, pair_tables=True, p2p=None):
        """lr: one float, or one per parameter (the reference's second parameter group, trainers/base.py:210-212).

        world_size > 1 makes step() own the gradient exchange of the data-parallel step:
        * small tensors (the MLPs): one flattened all-reduce, replicated update;
        * big tensors (the hash tables), under AMP: REDUCE-SCATTER of the gradient, Adam / EMA on this rank's 1/world
          shard only (optimizer state is allocated for the shard), then ALL-GATHER of the fp16 table copy -- the only
          form the next forward reads.  That moves 3/4 of an all-reduce's bytes and divides the optimizer pass by world.
          The fp32 master table is then current only inside each rank's shard (`master_complete` is False until
          gather_master() rebuilds it; state_dict() of the host-mirror modules does that); without AMP the tables are
          read in fp32 and the big tensors fall back to all-reduce.
        Which tensors are sharded / paired / exchanged depends only on (shape, requires_grad, world_size), never on the
        rank or on which gradients happen to exist, so every rank issues the same collectives.

        pair_tables (under AMP): when two big tensors share one [T, 2] shape (the model's two hash tables),
        their fp16 copies live in ONE interleaved buffer [T][table][2] and their gradients are accumulated by the
        paired scatter kernel into ONE interleaved f32 buffer (`grad_pair`) that this optimizer reads directly -- the
        tables' `.grad` stays None (use grad_of(p) to look at a gradient).  A corner of both tables is then one
        vector gather / one 16-byte reduction in the hash-grid kernels (nrf_grid_encode_forward_pair / _backward_pair).

        p2p (None = on when possible; env NRF_P2P=0 turns it off): with sharded paired tables on CUDA the whole exchange is
        fused into the optimizer kernels over NVLink / NVSwitch peer memory (csrc/optim_p2p.cu) -- the interleaved gradient
        and fp16 buffers become torch symmetric-memory allocations, each rank reduces its row shard straight out of every
        rank's gradient buffer (multimem.ld_reduce through the NVLS multicast mapping when the fabric has one), runs Adam /
        EMA on it and stores the fp16 rows into every rank's table copy (multimem.st); the MLP gradients and the found-inf
        flag go through a third symmetric buffer.  No NCCL call remains in the step; two symmetric-memory barriers bracket the
        kernel.  Anything else (no pairing, gloo, no symmetric memory) keeps the NCCL path."""
        self.params = [p for p in params]
        dev = self.params[0].device
        for p in self.params:                       # a previous fused optimizer over these parameters lets go of them
            prev = optimizer_of(p)
            if prev is not None:
                prev.detach()
        self.alive = True
        self.world_size, self.rank = int(world_size), int(rank)
        self.lrs = [float(v) for v in lr] if isinstance(lr, (list, tuple)) else [float(lr)] * len(self.params)
        if len(self.lrs) != len(self.params):
            raise ValueError('FusedAdamEMA: one learning rate per parameter expected')
        self.lr = self.lrs[0]
        self.betas, self.eps, self.lr_decay_steps = betas, eps, float(lr_decay_steps)
        self.ema_decay = ema_decay
        self.growth_factor, self.backoff_factor, self.growth_interval = growth_factor, backoff_factor, growth_interval
        self.enable_amp = enable_amp
        self.master_complete = True
        # shard plan: (lo, hi) element range of this rank for every sharded tensor, None for replicated ones.  Shards are
        # an even number of elements so that a [T, 2] row never straddles two ranks.
        self.shard = []
        for p in self.params:
            n = p.numel()
            if (self.world_size > 1 and shard_big and enable_amp and n >= half_copy_min_numel and n % self.world_size == 0
                    and (n // self.world_size) % 2 == 0):
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
                if (self.shard[big[0]] is None) == (self.shard[big[1]] is None) and self.lrs[big[0]] == self.lrs[big[1]]:
                    self.pair_idx = tuple(big)
        self.grad_pair = None           # [T, 2, 2] f32, allocated by the first paired backward
        self.grad_pair_valid = False    # holds this step's gradients (zero_grad() invalidates; the next backward clears it)
        self.grad_shard_pair = None
        # overlap_gather: the all-gather of the fp16 table copies runs on a side stream and the NEXT step's first table
        # read waits for it (current_half_copy) -- ray generation, near/far and the occupancy march of that step, which
        # do not touch the tables, overlap the exchange
        self.overlap_gather = True
        self.time_comm = False
        self.comm_events = []
        self._comm_stream = None
        self._gather_done = None
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
        me = weakref.ref(self)
        if self.pair_idx is not None:
            a, b = (self.params[i] for i in self.pair_idx)
            self.half_pair = torch.stack([a.detach().to(torch.float16), b.detach().to(torch.float16)], dim=1).contiguous()
        for i, p in enumerate(self.params):
            p._nrf_optimizer = me
            if self.pair_idx is not None and i in self.pair_idx:
                e = self.pair_idx.index(i)
                h = self.half_pair[:, e]                 # strided view: the interleaved buffer is the only fp16 copy
                p._nrf_half_copy = h
                p._nrf_half_version = p._version
                p._nrf_half_pair = (self.half_pair, e)
                p._nrf_grad_sink = (self, e)
                self.half.append(h)
                continue
            # big tensors (hash tables): fp16 copy only under AMP; small ones (the MLPs' flat params): always -- the
            # tensor-core kernels consume fp16 weights in every mode (nerfstyle_b200.tcnn.half_params)
            if (enable_amp and p.numel() >= half_copy_min_numel) or p.numel() < half_copy_min_numel:
                h = p.detach().to(torch.float16)
                p._nrf_half_copy = h
                p._nrf_half_version = p._version
                self.half.append(h)
            else:
                self.half.append(None)
        self._p2p, self._p2p_error = None, None
        if (p2p is None or p2p) and self.world_size > 1 and self.pair_idx is not None and self.shard[self.pair_idx[0]] is not None:
            if self._setup_p2p(dev):            # the interleaved buffers moved: re-register the views on the parameters
                for e, i in enumerate(self.pair_idx):
                    p = self.params[i]
                    h = self.half_pair[:, e]
                    p._nrf_half_copy, p._nrf_half_pair = h, (self.half_pair, e)
                    self.half[i] = h
            elif p2p:
                raise RuntimeError('FusedAdamEMA(p2p=True): symmetric memory unavailable (%s)' % self._p2p_error)

    # ------------------------------------------------------------------------------------------------ peer memory
    def _setup_p2p(self, dev):
        """Re-home the interleaved gradient / fp16 buffers in symmetric memory and rendezvous (a COLLECTIVE: every rank
        constructs its optimizer at the same point).  Returns False (leaving everything as it was) when unavailable."""
        import os
        import torch.distributed as dist
        if os.environ.get('NRF_P2P', '1') == '0' or dev.type != 'cuda' or not dist.is_initialized() or dist.get_backend() != 'nccl':
            return False
        try:
            import torch.distributed._symmetric_memory as symm
            group = dist.group.WORLD
            ia, ib = self.pair_idx
            T = self.params[ia].shape[0]
            self._small_idx = [i for i, sh in enumerate(self.shard) if sh is None and (self.pair_idx is None or i not in self.pair_idx)]
            n_small = sum(self.params[i].numel() for i in self._small_idx)
            gp = symm.empty((T, 2, 2), dtype=torch.float32, device=dev)
            hp = symm.empty((T, 2, 2), dtype=torch.float16, device=dev)
            sm = symm.empty((n_small + 8,), dtype=torch.float32, device=dev)
            gp.zero_(); sm.zero_()
            hp.copy_(self.half_pair)
            hg, hh, hs = symm.rendezvous(gp, group), symm.rendezvous(hp, group), symm.rendezvous(sm, group)
            as_dev = lambda h: torch.tensor([int(p) for p in h.buffer_ptrs], dtype=torch.int64, device=dev)      # noqa: E731
            self._p2p = {'hg': hg, 'hh': hh, 'hs': hs, 'grad_ptrs': as_dev(hg), 'half_ptrs': as_dev(hh), 'small_ptrs': as_dev(hs),
                         'grad_mc': int(hg.multicast_ptr) if hg.has_multicast_support else 0,
                         'half_mc': int(hh.multicast_ptr) if hh.has_multicast_support else 0,
                         'small': sm, 'n_small': n_small, 'small_red': torch.zeros(n_small, dtype=torch.float32, device=dev)}
            if os.environ.get('NRF_P2P_NO_MULTICAST', '0') == '1':
                self._p2p['grad_mc'] = self._p2p['half_mc'] = 0
        except Exception as e:        # symmetric memory not available on this build / fabric: NCCL path
            self._p2p = None
            self._p2p_error = '%s: %s' % (type(e).__name__, e)
            return False
        self.grad_pair, self.half_pair = gp, hp
        self.grad_pair_valid = False
        return True

    def _step_p2p(self, grads, lib, dev, st, omd):
        """The N > 1 step with the exchange fused into the kernels (see __init__ / csrc/optim_p2p.cu)."""
        P = self._p2p
        ia, ib = self.pair_idx
        sh = self.shard[ia]
        # 1. local inf / nan check BEFORE the exchange (a non-finite sum needs a non-finite addend), flag + small gradients
        #    into this rank's symmetric buffer
        L.check(lib.nrf_grads_check(self.grad_pair.data_ptr(), self.grad_pair.numel(), self.state.data_ptr(), st), 'grads_check')
        o = 0
        for i in self._small_idx:
            g, n = grads[i], self.params[i].numel()
            if g is None:
                P['small'][o:o + n].zero_()
            else:
                L.check(lib.nrf_grads_check(g.data_ptr(), n, self.state.data_ptr(), st), 'grads_check')
                P['small'][o:o + n].copy_(g.reshape(-1))
            o += n
        others = [i for i, s_ in enumerate(self.shard) if s_ is not None and i not in self.pair_idx]      # sharded, not paired
        for i in others:
            if grads[i] is None and self.params[i].requires_grad:
                grads[i] = torch.zeros(self.params[i].shape, dtype=torch.float32, device=dev)
            if grads[i] is not None:
                L.check(lib.nrf_grads_check(grads[i].data_ptr(), grads[i].numel(), self.state.data_ptr(), st), 'grads_check')
        P['small'][o:o + 1].copy_(self.state[4:8].view(torch.int32))
        # 2. every rank's gradients are complete
        with self._timed('barrier'):
            P['hg'].barrier()
        # 3. small tensors + the skip decision, identical on every rank (summed in rank order)
        with self._timed('p2p_small'):
            L.check(lib.nrf_small_allreduce_p2p(P['small_ptrs'].data_ptr(), self.world_size, P['n_small'], P['small_red'].data_ptr(),
                                                self.state.data_ptr(), st), 'small_allreduce_p2p')
        # 4. reduce-scatter + Adam / EMA + all-gather of the paired tables in one kernel over peer memory
        lo, n = sh[0], sh[1] - sh[0]
        with self._timed('p2p_adam_pair'):
            L.check(lib.nrf_adam_step_pair_p2p(
                self.params[ia].data_ptr() + 4 * lo, self.params[ib].data_ptr() + 4 * lo, P['grad_ptrs'].data_ptr(), P['grad_mc'] or None,
                P['half_ptrs'].data_ptr(), P['half_mc'] or None, self.world_size, lo // 2,
                self.exp_avg[ia].data_ptr(), self.exp_avg[ib].data_ptr(), self.exp_avg_sq[ia].data_ptr(), self.exp_avg_sq[ib].data_ptr(),
                L.ptr(self.ema[ia]) if self.ema is not None else None, L.ptr(self.ema[ib]) if self.ema is not None else None, n // 2,
                self.state.data_ptr(), self.lrs[ia], self.lr_decay_steps, self.betas[0], self.betas[1], self.eps, omd, st),
                'adam_step_pair_p2p')
        self.master_complete = False
        # 5. the replicated small tensors from the reduced gradients
        o = 0
        for i in self._small_idx:
            p, n = self.params[i], self.params[i].numel()
            if p.requires_grad:
                h_ptr = self.half[i].data_ptr() if self.half[i] is not None else None
                L.check(lib.nrf_adam_step_ex(p.data_ptr(), P['small_red'].data_ptr() + 4 * o, self.exp_avg[i].data_ptr(),
                                             self.exp_avg_sq[i].data_ptr(), L.ptr(self.ema[i]) if self.ema is not None else None, h_ptr, n,
                                             self.state.data_ptr(), self.lrs[i], self.lr_decay_steps, self.betas[0], self.betas[1],
                                             self.eps, omd, 1, 1, st), 'adam_step')
            o += n
        # 5b. sharded tensors outside the pair (none in the reference's model) keep the NCCL exchange
        if others:
            from . import parallel
            for i in others:
                if grads[i] is None:
                    continue
                so = self.shard[i]
                parallel.reduce_scatter_sum(grads[i].view(-1), self.grad_shard[i], self.world_size, self.rank)
                h_ptr = (self.half[i].data_ptr() + 2 * so[0]) if self.half[i] is not None else None
                L.check(lib.nrf_adam_step_ex(self.params[i].data_ptr() + 4 * so[0], self.grad_shard[i].data_ptr(), self.exp_avg[i].data_ptr(),
                                             self.exp_avg_sq[i].data_ptr(), L.ptr(self.ema[i]) if self.ema is not None else None, h_ptr,
                                             so[1] - so[0], self.state.data_ptr(), self.lrs[i], self.lr_decay_steps, self.betas[0],
                                             self.betas[1], self.eps, omd, 1, 1, st), 'adam_step')
                if self.half[i] is not None:
                    hflat = self.half[i].view(-1)
                    parallel.all_gather_shards(hflat, hflat[so[0]:so[1]], self.world_size)
        # 6. every rank's table rows have landed (and nobody still reads this rank's gradients): on the side stream, the next
        #    step's first table read waits for it
        side = None
        if self.overlap_gather:
            if self._comm_stream is None:
                self._comm_stream = torch.cuda.Stream(device=dev)
            side = self._comm_stream
            side.wait_stream(torch.cuda.current_stream(dev))
        with (torch.cuda.stream(side) if side is not None else contextlib.nullcontext()):
            with self._timed('barrier_tail'):
                P['hh'].barrier()
            if side is not None:
                self._gather_done = torch.cuda.Event()
                self._gather_done.record(side)

    # ------------------------------------------------------------------------------------------------ ownership
    def detach(self):
        """Let go of the parameters: remove every attribute this optimizer hung on them (fp16 shadows, pair buffer,
        gradient sink).  The forward then casts the tables itself and the backward produces ordinary `.grad`s again --
        e.g. before handing the model to another optimizer (the reference's _reset_optim between training stages)."""
        for p in self.params:
            ref = getattr(p, '_nrf_optimizer', None)
            if ref is not None and ref() is self:
                for a in _ATTRS:
                    if hasattr(p, a):
                        delattr(p, a)
        self.alive = False

    close = detach

    def wait_pending_gather(self):
        """Make the current stream wait for an all-gather still in flight on the side stream (no host sync)."""
        ev = self._gather_done
        if ev is not None:
            with self._timed('wait_gather'):
                torch.cuda.current_stream(self.params[0].device).wait_event(ev)
            self._gather_done = None

    @contextlib.contextmanager
    def _timed(self, name):
        """bench.py only (time_comm = True): CUDA events around a collective / a wait on the stream it is issued on."""
        if not self.time_comm or self.params[0].device.type != 'cuda':
            yield
            return
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        yield
        e1.record()
        self.comm_events.append((name, e0, e1))

    def _require_alive(self):
        if not self.alive:
            raise RuntimeError('FusedAdamEMA: this optimizer was detached from its parameters (a newer one owns them)')

    def is_sharded(self):
        return any(sh is not None for sh in self.shard)

    @torch.no_grad()
    def refresh_half_copies(self, only=None, external_write=False):
        """Re-cast the fp16 shadows from the fp32 parameters: after writing the parameters by any means PyTorch's version
        counter does not see (`param.data` writes, e.g. torch_ema's copy_to / restore).  Version-visible writes
        (load_state_dict, param.copy_) are picked up automatically by the next forward.  With a sharded optimizer the
        fp32 master is complete only after gather_master() -- unless the write being reported replaced the whole tensor
        (external_write=True, what the automatic path passes)."""
        self._require_alive()
        self.wait_pending_gather()
        if self.is_sharded() and not self.master_complete and not external_write:
            self.gather_master()
        for p, h in zip(self.params, self.half):
            if h is not None and (only is None or p is only):
                h.copy_(p.detach())
                p._nrf_half_version = p._version
        if external_write and self.is_sharded():
            # a whole-tensor write from outside (checkpoint load) makes the master complete again; the optimizer shards
            # of EMA keep their own history
            if only is None or all(p._version == p._nrf_half_version for p, h in zip(self.params, self.half) if h is not None):
                self.master_complete = True

    def scale_loss(self, loss):
        return loss * self.scale if self.enable_amp else loss

    def zero_grad(self):
        for p in self.params:
            p.grad = None
        self.grad_pair_valid = False

    def grad_pair_buffer(self):
        """The interleaved gradient buffer the paired scatter accumulates into (called by the dual encoder's backward)."""
        self._require_alive()
        self.wait_pending_gather()       # peer ranks may still be reading last step's gradients (peer-memory exchange)
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

    # ------------------------------------------------------------------------------------------------ sharded state
    @torch.no_grad()
    def gather_master(self):
        """Rebuild the full fp32 master copy of every sharded tensor on every rank (checkpointing; a COLLECTIVE)."""
        from . import parallel
        for p, sh in zip(self.params, self.shard):
            if sh is not None:
                flat = p.data.view(-1)
                parallel.all_gather_shards(flat, flat[sh[0]:sh[1]].clone(), self.world_size)
        self.master_complete = True

    @torch.no_grad()
    def full_ema(self):
        """The EMA weights as full tensors shaped like the parameters (sharded ones are all-gathered: a COLLECTIVE)."""
        from . import parallel
        if self.ema is None:
            return None
        out = []
        for p, e, sh in zip(self.params, self.ema, self.shard):
            if sh is None:
                out.append(e)
            else:
                full = torch.empty_like(p).view(-1)
                parallel.all_gather_shards(full, e.contiguous(), self.world_size)
                out.append(full.view_as(p))
        return out

    @contextlib.contextmanager
    def ema_scope(self):
        """torch_ema's `average_parameters()`: inside the scope the model evaluates with the EMA weights (the reference
        tests / renders under EMA, trainers/base.py:361-377), fp16 shadows included; the training weights come back on
        exit.  With a sharded optimizer entering and leaving are collectives."""
        self._require_alive()
        if self.ema is None:
            yield
            return
        with torch.no_grad():
            if self.is_sharded() and not self.master_complete:
                self.gather_master()
            backup = [p.detach().clone() for p in self.params]
            for p, e in zip(self.params, self.full_ema()):
                p.data.copy_(e)
            self.refresh_half_copies()
        try:
            yield
        finally:
            with torch.no_grad():
                for p, b in zip(self.params, backup):
                    p.data.copy_(b)
                self.refresh_half_copies()

    def _gather_half_copies(self, dev, has_grad, pair_grads):
        """All-gather of the fp16 copies of the sharded tensors that were just updated.  With overlap_gather it is issued
        on a side stream (after this step's Adam kernels) and `_gather_done` is recorded for the next table read."""
        from . import parallel
        todo = [i for i, sh in enumerate(self.shard)
                if sh is not None and (has_grad[i] or (pair_grads and i in self.pair_idx))]
        if not todo:
            return
        self.master_complete = False
        side = None
        if self.overlap_gather and dev.type == 'cuda':
            if self._comm_stream is None:
                self._comm_stream = torch.cuda.Stream(device=dev)
            side = self._comm_stream
            side.wait_stream(torch.cuda.current_stream(dev))
        with (torch.cuda.stream(side) if side is not None else contextlib.nullcontext()):
            gathered_pair = False
            for i in todo:
                sh = self.shard[i]
                if self.pair_idx is not None and i in self.pair_idx:
                    if not gathered_pair:          # param elements [lo, hi) <-> interleaved halfs [2 lo, 2 hi)
                        hflat = self.half_pair.view(-1)
                        with self._timed('all_gather'):
                            parallel.all_gather_shards(hflat, hflat[2 * sh[0]:2 * sh[1]], self.world_size)
                        gathered_pair = True
                    continue
                hflat = self.half[i].view(-1)
                with self._timed('all_gather'):
                    parallel.all_gather_shards(hflat, hflat[sh[0]:sh[1]], self.world_size)
            if side is not None:
                self._gather_done = torch.cuda.Event()
                self._gather_done.record(side)

    # ------------------------------------------------------------------------------------------------ the step
    @torch.no_grad()
    def step(self):
        from . import parallel
        self._require_alive()
        self.wait_pending_gather()          # (a step without a forward in between: tests)
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
            pair_trainable = self.pair_idx is not None and all(self.params[i].requires_grad for i in self.pair_idx)
            pair_has_grad = self.pair_idx is not None and any(grads[i] is not None for i in self.pair_idx)
            if pair_has_grad and self.grad_pair_valid:
                raise RuntimeError('FusedAdamEMA.step: a paired table has BOTH an ordinary .grad and gradients in the interleaved '
                                   'pair buffer (two backward paths were mixed in one step); call zero_grad() between them')
            pair_grads = pair_trainable and not pair_has_grad
            if pair_grads and not self.grad_pair_valid and (self.world_size > 1 or self.grad_pair is not None):
                self.grad_pair_buffer()          # this rank produced no table gradient (empty batch): exchange zeros
            pair_grads = pair_grads and self.grad_pair_valid
            pair_src = self.grad_pair if pair_grads else None      # tensor holding this rank's interleaved gradient rows
            if self._p2p is not None and pair_grads:
                self._step_p2p(grads, lib, dev, st, omd)
                L.check(lib.nrf_scaler_update(self.state.data_ptr(), self.growth_factor, self.backoff_factor,
                                              int(self.growth_interval) if self.enable_amp else (1 << 30), st), 'scaler_update')
                return
            if self.world_size > 1:
                # every rank must issue the same collectives: a trainable tensor without a gradient contributes zeros
                for i, p in enumerate(self.params):
                    if grads[i] is None and p.requires_grad and not (pair_grads and i in self.pair_idx):
                        grads[i] = torch.zeros(p.shape, dtype=torch.float32, device=dev)
                # ---- gradient exchange: reduce-scatter the sharded tensors, one flattened all-reduce for the rest
                for i, sh in enumerate(self.shard):
                    if sh is not None and grads[i] is not None:
                        with self._timed('reduce_scatter'):
                            parallel.reduce_scatter_sum(grads[i].view(-1), self.grad_shard[i], self.world_size, self.rank)
                        grads[i] = self.grad_shard[i]
                if pair_grads:
                    if self.shard[self.pair_idx[0]] is not None:
                        # row shards of the two tables coincide: ONE reduce-scatter of the interleaved buffer
                        flat = self.grad_pair.view(-1)
                        if self.grad_shard_pair is None:
                            self.grad_shard_pair = torch.empty(flat.numel() // self.world_size, dtype=torch.float32, device=dev)
                        with self._timed('reduce_scatter'):
                            parallel.reduce_scatter_sum(flat, self.grad_shard_pair, self.world_size, self.rank)
                        pair_src = self.grad_shard_pair
                    else:
                        with self._timed('all_reduce'):
                            parallel.allreduce_tensors([self.grad_pair], self.world_size)
                with self._timed('all_reduce_small'):
                    parallel.allreduce_tensors([g for g, sh in zip(grads, self.shard) if sh is None and g is not None], self.world_size)
            for g in grads:
                if g is not None:
                    L.check(lib.nrf_grads_check(g.data_ptr(), g.numel(), self.state.data_ptr(), st), 'grads_check')
            if pair_grads:
                L.check(lib.nrf_grads_check(pair_src.data_ptr(), pair_src.numel(), self.state.data_ptr(), st), 'grads_check')
            if self.world_size > 1 and self.is_sharded():
                # an inf seen in ANY rank's shard skips the step everywhere (the replicated GradScaler decision)
                with self._timed('all_reduce_small'):
                    parallel.allreduce_max_int(self.state[4:8].view(torch.int32), self.world_size)
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
                    self.state.data_ptr(), self.lrs[ia], self.lr_decay_steps, self.betas[0], self.betas[1], self.eps, omd, st),
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
                                             self.state.data_ptr(), self.lrs[i], self.lr_decay_steps, self.betas[0], self.betas[1],
                                             self.eps, omd, g_stride, h_stride, st), 'adam_step')
            if self.world_size > 1:
                # ---- the next forward reads only the fp16 table copies: gather their shards (in place)
                self._gather_half_copies(dev, [g is not None for g in grads], pair_grads)
            L.check(lib.nrf_scaler_update(self.state.data_ptr(), self.growth_factor, self.backoff_factor,
                                          int(self.growth_interval) if self.enable_amp else (1 << 30), st), 'scaler_update')
