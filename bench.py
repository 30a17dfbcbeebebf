#!/usr/bin/env python
"""bench.py -- the reference's headline metric on B200: train-step rays/s of the NeRF hot path.

Workload (BASELINE.json configs[1]): LLFF-room-shaped reconstruction train step, 504x378 views, 8192 rays per
step per GPU, occupancy-grid marching (H=128, 2 cascades, bound 2), two 16-level 2^19 hash grids, four 64-wide
MLPs (K=8 classes), composite fwd+bwd, MSE + class CE, AMP (fp16 tables/MLPs, GradScaler), Adam + EMA,
occupancy update every 16 steps; random-init field, synthetic poses/targets (no dataset on the box).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]

`value`  : rays/s with the step's inputs already resident in HBM (device timed, max over ranks).
`e2e`    : same metric through the public API with HOST (pinned) inputs: H2D of rays+targets and D2H of the loss
           inside the timed region, every step.
`roofline`: the dominant kernel of the step (decided from CUDA-event timings taken live over the timed region).
`cpu_baseline`: the CPU oracle (port of the reference kernels, OpenMP + torch CPU) on a bounded sample.
--impl reference: the reference path has no CPU implementation of its own (CUDA-only), so the arm times the
           oracle port on all host cores on bounded samples of the same workload.
"""
import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

N_RAYS = 8192
N_CLASSES = 8
WORKLOAD = 'llff_room_train_step_8192rays_504x378'


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument('--gpus', type=int, default=1)
    ap.add_argument('--steps', type=int, default=64)
    ap.add_argument('--warmup', type=int, default=8)
    ap.add_argument('--impl', default='ours', choices=['ours', 'reference'])
    ap.add_argument('--rays', type=int, default=N_RAYS)
    ap.add_argument('--no-amp', action='store_true')
    ap.add_argument('--cpu-sample-rays', type=int, default=1024)
    ap.add_argument('--skip-cpu-baseline', action='store_true')
    ap.add_argument('--skip-ref-ext', action='store_true')
    ap.add_argument('--skip-extras', action='store_true')
    ap.add_argument('--skip-strong', action='store_true')
    return ap.parse_args()


# ------------------------------------------------------------------------------------------------ clocks
class ClockSampler:
    """SM clock / throttle reasons sampled DURING the timed region.  In-process NVML (pynvml) on a background thread:
    spawning `nvidia-smi` every 200 ms costs ~1 s of CPU per call and takes driver locks that stall kernel launches
    (it made the timed region itself 10-100 % slower); nvidia-smi is only the fallback when pynvml is missing."""
    NAMES = {0x8: 'hw_slowdown', 0x40: 'hw_thermal_slowdown', 0x20: 'sw_thermal_slowdown', 0x4: 'sw_power_cap'}

    def __init__(self, index=0, period=0.1):
        self.index, self.period, self.sm, self.reasons, self.power = index, period, [], set(), []
        self.max_mhz, self._stop, self._t, self.h, self.nv = None, threading.Event(), None, None, None
        try:
            import pynvml
            import torch
            pynvml.nvmlInit()
            pr = torch.cuda.get_device_properties(index)
            bus = '%08x:%02x:%02x.0' % (pr.pci_domain_id, pr.pci_bus_id, pr.pci_device_id)
            try:
                self.h = pynvml.nvmlDeviceGetHandleByPciBusId(bus.encode())
            except Exception:
                self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.nv = pynvml
            self.max_mhz = float(pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM))
        except Exception:
            self.nv = None
        if self.nv is not None:
            # the FIRST query of each NVML counter initialises driver state (tens of ms, under a lock that stalls kernel
            # launches): pay for it here, outside the timed region, and drop the values
            try:
                self._sample()
            except Exception:
                pass
            self.sm, self.reasons, self.power = [], set(), []

    def _sample(self):
        if self.nv is not None:
            nv = self.nv
            self.sm.append(float(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM)))
            try:
                self.power.append(nv.nvmlDeviceGetPowerUsage(self.h) / 1000.0)
            except Exception:
                pass
            try:
                mask = nv.nvmlDeviceGetCurrentClocksEventReasons(self.h)
            except Exception:
                mask = nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
            for bit, name in self.NAMES.items():
                if mask & bit:
                    self.reasons.add(name)
        else:
            q = 'clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,' \
                'clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap'
            out = subprocess.run(['nvidia-smi', '-i', str(self.index), '--query-gpu=' + q, '--format=csv,noheader,nounits'],
                                 capture_output=True, text=True, timeout=5).stdout.strip().split(',')
            self.sm.append(float(out[0]))
            self.max_mhz = float(out[1])
            for n, v in zip(['hw_slowdown', 'hw_thermal_slowdown', 'sw_thermal_slowdown', 'sw_power_cap'], out[2:6]):
                if v.strip().lower().startswith('active'):
                    self.reasons.add(n)

    def _run(self):
        while not self._stop.is_set():
            try:
                self._sample()
            except Exception:
                pass
            self._stop.wait(self.period if self.nv is not None else 1.0)

    def start(self):
        self._t = threading.Thread(target=self._run, daemon=True)
        self._t.start()

    def stop(self):
        self._stop.set()
        if self._t:
            self._t.join(timeout=6)
        return {'sm_mhz': statistics.median(self.sm) if self.sm else None, 'sm_max_mhz': self.max_mhz,
                'reasons': sorted(self.reasons), 'samples': len(self.sm),
                'power_w_max': round(max(self.power), 1) if self.power else None,
                'source': 'pynvml' if self.nv is not None else 'nvidia-smi'}


# ------------------------------------------------------------------------------------------------ workload
def make_batches(n_steps, n_rays, rank, world, device):
    """Per-step (rays_o, rays_d, target_rgb, target_cls) for this rank, generated once; returned both as pinned host
    tensors (e2e) and as device tensors (device-resident value)."""
    import torch
    from nerfstyle_b200 import scenes
    intr = dict(scenes.ROOM)
    poses = scenes.synthetic_poses(intr['n_train'], 0)
    gen = torch.Generator().manual_seed(69420 + 1000 * rank)      # rng_seed of cfgs/training/default.yaml
    host, dev = [], []
    for s in range(n_steps):
        pose = poses[(s * world + rank) % len(poses)]
        idx = scenes.frame_indices(intr, n_rays, gen).to(device)
        o, d = scenes.generate_rays(pose, intr, device, idx)
        rgb, seg = scenes.synthetic_target(idx, intr, N_CLASSES)
        pack = torch.cat([o, d, rgb, seg.to(torch.float32)[:, None]], dim=1).contiguous()      # [n, 10] f32
        dev.append(pack)
        host.append(pack.cpu().pin_memory())
    return host, dev


def unpack(pack):
    return pack[:, 0:3].contiguous(), pack[:, 3:6].contiguous(), pack[:, 6:9].contiguous(), pack[:, 9].long()


def build_trainer(device, amp, world):
    import torch
    from nerfstyle_b200 import model as M
    from nerfstyle_b200.trainer import TrainStep
    torch.manual_seed(0)
    m = M.StyleTCNerf([-2., -2., -2.], [2., 2., 2.], class_dim=N_CLASSES).to(device)
    r = M.Renderer(m, 2.0, raymarch_channels=3 + N_CLASSES).to(device)
    return TrainStep(r, enable_amp=amp, world_size=world)


ALGO_BYTES = {   # SURVEY.md 8d, per point per encoder (fp32 tables / fp16 tables), per sample for compositing
    'nrf_grid_encode_forward': {False: 12 + 16 * 8 * 2 * 4 + 16 * 2 * 4, True: 12 + 16 * 8 * 2 * 2 + 16 * 2 * 2},
    'nrf_grid_encode_backward': {False: 12 + 128 + 2 * 1024, True: 12 + 64 + 2 * 512},
    # one pass over the points for TWO encoders: the xyz read is shared
    'nrf_grid_encode_forward_dual': {False: 12 + 2 * (16 * 8 * 2 * 4 + 16 * 2 * 4), True: 12 + 2 * (16 * 8 * 2 * 2 + 16 * 2 * 2)},
    'nrf_grid_encode_backward_dual': {False: 12 + 2 * (128 + 2 * 1024), True: 12 + 2 * (64 + 2 * 512)},
    # paired (interleaved-table) forms: same algorithmic bytes as the dual forms, half the gathers / reductions
    'nrf_grid_encode_forward_pair': {False: 12 + 2 * (16 * 8 * 2 * 4 + 16 * 2 * 4), True: 12 + 2 * (16 * 8 * 2 * 2 + 16 * 2 * 2)},
    'nrf_grid_encode_backward_pair': {False: 12 + 2 * (128 + 2 * 1024), True: 12 + 2 * (64 + 2 * 512)},
}


def run_ours(args):
    import torch
    import torch.distributed as dist
    from nerfstyle_b200 import _lib
    world = int(os.environ.get('WORLD_SIZE', '1'))
    rank = int(os.environ.get('RANK', '0'))
    local = int(os.environ.get('LOCAL_RANK', '0'))
    torch.cuda.set_device(local)
    device = torch.device('cuda', local)
    if world > 1:
        dist.init_process_group('nccl', device_id=device)
    amp = not args.no_amp
    _lib.lib()
    W, K = args.warmup, args.steps
    host, devb = make_batches(W + K, args.rays, rank, world, device)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def fresh_trainer():
        """Both timed phases start from the SAME initial state (same seeds, same batches) so that `value` and `e2e`
        run the same training trajectory (the sample count per step drifts as the density field trains)."""
        torch.manual_seed(0)
        torch.cuda.manual_seed_all(0)
        return build_trainer(device, amp, world)

    # -------- phase A: device-resident inputs (value) + live per-kernel CUDA-event timing
    from nerfstyle_b200.trainer import TrainStep
    TrainStep.reserve_workspace(device)      # set-up, not a step: one arena instead of cudaMallocs at every new high-water mark
    ts = fresh_trainer()
    for s in range(W):
        ts.step(*unpack(devb[s]))
    # live CUDA-event timing inside the timed region: only the roofline candidates (4 calls per step); the full per-op
    # breakdown (`kernels`) is taken in a separate instrumented pass afterwards so that it does not perturb `value`
    all_ops = ['nrf_grid_encode_forward', 'nrf_grid_encode_backward', 'nrf_grid_encode_forward_dual',
               'nrf_grid_encode_backward_dual', 'nrf_grid_encode_forward_pair', 'nrf_grid_encode_backward_pair', 'nrf_mlp_forward', 'nrf_mlp_backward',
               'nrf_mlp_forward_ex', 'nrf_mlp_backward_ex', 'nrf_field_forward',
               'nrf_composite_rays_train_forward', 'nrf_composite_rays_train_backward', 'nrf_composite_rays_train_backward_ex',
               'nrf_march_rays_train_count', 'nrf_march_rays_train_count_staged', 'nrf_march_rays_train_emit',
               'nrf_march_rays_train_write']
    timed_ops = ['nrf_grid_encode_forward', 'nrf_grid_encode_backward', 'nrf_grid_encode_forward_dual',
                 'nrf_grid_encode_backward_dual', 'nrf_grid_encode_forward_pair', 'nrf_grid_encode_backward_pair']
    barrier()
    clocks = ClockSampler(local)
    clocks.start()
    _lib.Stats.reset()
    _lib.Stats.timed = set() if os.environ.get('NRF_BENCH_NO_EVENTS') == '1' else set(timed_ops)
    seg0 = torch.cuda.memory_stats(device).get('segment.all.allocated', 0)
    barrier()
    e0 = torch.cuda.Event(enable_timing=True)
    e1 = torch.cuda.Event(enable_timing=True)
    e0.record()
    for s in range(W, W + K):
        loss = ts.step(*unpack(devb[s]))
    e1.record()
    barrier()
    ms_total = e0.elapsed_time(e1)
    seg1 = torch.cuda.memory_stats(device).get('segment.all.allocated', 0)
    launches = _lib.Stats.launches
    events = list(_lib.Stats.events)
    _lib.Stats.timed = set()
    t = torch.tensor([ms_total], device=device, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_total = float(t.item())
    per_op = {}
    for name, a, b, units in events:
        d = per_op.setdefault(name, {'ms': 0.0, 'calls': 0, 'units': 0})
        d['ms'] += a.elapsed_time(b)
        d['calls'] += 1
        d['units'] += units
    samples_last = int(ts.renderer.step_counter[(ts.renderer.local_step - 1) % 16, 0].item())
    # separate instrumented pass (8 steps) for the per-op breakdown
    _lib.Stats.reset()
    _lib.Stats.timed = set(all_ops)
    nb = min(8, W + K)
    if ts.fused is not None:
        ts.fused.time_comm, ts.fused.comm_events = world > 1, []
    for s in range(nb):
        ts.step(*unpack(devb[s]))
    torch.cuda.synchronize()
    comm = {}

    class ts_p2p:
        value = ts.fused is not None and ts.fused._p2p is not None
        multicast = bool(value and ts.fused._p2p['grad_mc'])
    if ts.fused is not None and world > 1:
        for name, a, b in ts.fused.comm_events:
            comm[name] = comm.get(name, 0.0) + a.elapsed_time(b) / nb
        ts.fused.time_comm = False
    breakdown = {}
    for name, a, b, units in _lib.Stats.events:
        d = breakdown.setdefault(name, {'ms': 0.0, 'calls': 0})
        d['ms'] += a.elapsed_time(b)
        d['calls'] += 1
    _lib.Stats.timed = set()
    _lib.Stats.events = []
    del ts      # the caching allocator stays warm: steady-state training does not re-cudaMalloc its sample buffers

    # -------- phase B: end to end through the public API with host inputs (H2D + D2H every step)
    ts = fresh_trainer()
    stage = [torch.empty_like(devb[0]), torch.empty_like(devb[0])]
    loss_pin = torch.zeros(1, dtype=torch.float32).pin_memory()
    for s in range(W):
        stage[s & 1].copy_(host[s], non_blocking=True)
        ts.step(*unpack(stage[s & 1]), loss_host=loss_pin)
        ts.loss_ready.synchronize()
        float(loss_pin[0])
    barrier()
    t0 = time.perf_counter()
    for s in range(W, W + K):
        stage[s & 1].copy_(host[s], non_blocking=True)                 # H2D from pinned memory (double-buffered staging)
        ts.step(*unpack(stage[s & 1]), loss_host=loss_pin)             # D2H of the loss on a side stream, after the forward
        ts.loss_ready.synchronize()
        lv = float(loss_pin[0])                                        # the step's loss, on the host, every step
    torch.cuda.synchronize()
    e2e_s = time.perf_counter() - t0
    t = torch.tensor([e2e_s], device=device, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    e2e_s = float(t.item())
    clk = clocks.stop()

    # -------- strong-scaling point: a FIXED global batch of 65 536 rays per step split over the ranks (the default line
    # above is weak scaling: 8192 rays per GPU).  Same trainer state, 2 warm-up + 6 timed steps, device-resident inputs.
    strong = None
    if not args.skip_strong:
        g_rays = 65536
        per = g_rays // world
        # ~520 samples per ray x ~700 B of activations per sample: give the caching allocator one arena of that size first
        # (set-up, like reserve_workspace above), so the timed steps do not cudaMalloc at each new high-water mark
        TrainStep.reserve_workspace(device, min(int(per * 520 * 900), 64 << 30))
        _, sb = make_batches(8, per, rank, world, device)
        for s in range(2):
            ts.step(*unpack(sb[s]))
        barrier()
        s0, s1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s0.record()
        for s in range(2, 8):
            ts.step(*unpack(sb[s]))
        s1.record()
        barrier()
        t = torch.tensor([s0.elapsed_time(s1)], device=device, dtype=torch.float64)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        sms = float(t.item()) / 6
        strong = {'global_rays_per_step': g_rays, 'rays_per_gpu_per_step': per, 'steps': 6, 'ms_per_step': round(sms, 4),
                  'value': round(g_rays / (sms / 1e3), 1), 'unit': 'rays/s', 'scaling': 'strong'}
        del sb

    # -------- full-frame inference at N > 1 (BASELINE config 3: row tiles sharded over the ranks, image gathered on rank 0)
    render_n = None
    if world > 1 and not args.skip_extras:
        render_n = render_sharded(device, rank, world)

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return
    n_global = args.rays * world
    value = n_global * K / (ms_total / 1e3)
    # dominant kernel -> roofline
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, 'MEASURED_PEAKS.json')))
    except Exception:
        pass
    hbm_peak = float(peaks.get('hbm_gbs', 6650.0))
    peak_src = 'measured (MEASURED_PEAKS.json hbm_gbs)' if 'hbm_gbs' in peaks else 'fallback 6650 GB/s (B200_PROFILING.md)'
    kern = {k: {'ms_per_step': round(v['ms'] / nb, 4), 'calls_per_step': v['calls'] / nb} for k, v in breakdown.items()}
    top = max((k for k in per_op if k in ALGO_BYTES), key=lambda k: per_op[k]['ms'], default=None)
    roofline = None
    traffic = {}
    try:      # dram__bytes_read.sum + dram__bytes_write.sum per launch from the committed `ncu --set full` capture
        traffic = json.load(open(os.path.join(ROOT, 'profiles', 'traffic.json')))
    except Exception:
        pass
    if top:
        v = per_op[top]
        bytes_per_launch = ALGO_BYTES[top][amp] * (v['units'] / v['calls'])
        sec_per_launch = v['ms'] / 1e3 / v['calls']
        ach = bytes_per_launch / sec_per_launch / 1e9
        roofline = {'kernel': top, 'bound': 'hbm', 'achieved': round(ach, 1), 'peak': hbm_peak, 'unit': 'GB/s',
                    'frac': round(ach / hbm_peak, 4), 'traffic': (traffic.get(top) or {}).get('dram_bytes_per_launch'),
                    'traffic_source': (traffic.get(top) or {}).get('source'), 'peak_source': peak_src,
                    'algorithmic_bytes_per_point': ALGO_BYTES[top][amp], 'points_per_launch': v['units'] / v['calls'],
                    'us_per_launch': round(sec_per_launch * 1e6, 1),
                    'hbm_frac_actual': (round((traffic.get(top) or {}).get('dram_bytes_per_launch') / sec_per_launch / 1e9 / hbm_peak, 4)
                                        if (traffic.get(top) or {}).get('dram_bytes_per_launch') else None),
                    'binding_counter': (traffic.get(top) or {}).get('binding_counter'),
                    'note': 'algorithmic bytes (SURVEY 8d) charge a read-modify-write of every table row a point touches; the walk-form '
                            'scatter sums consecutive samples of a ray in registers and touches a row once per cell visit, and both '
                            'gradient tables stay L2-resident, so real DRAM traffic (`traffic`) is far lower and frac exceeds 1 -- ncu '
                            'shows the kernel bound by the LSU data pipe and the L1->crossbar request rate of its reductions, '
                            'profiles/r02c_ncu_kernels.md'}
    line = {
        'metric': 'train_rays_per_s', 'value': round(value, 1), 'unit': 'rays/s', 'n_gpus': world, 'steps': K, 'warmup': W,
        'ms_per_step': round(ms_total / K, 4), 'higher_is_better': True, 'scaling': 'weak', 'vs_baseline': None,
        'dtype': 'f16 tables+MLP operands / f32 accumulate+compositing' if amp else 'f32 tables / f16 MLP operands',
        'data': 'synthetic',
        'config': {'workload': WORKLOAD, 'rays_per_gpu_per_step': args.rays, 'global_rays_per_step': n_global,
                   'samples_per_step_last': samples_last, 'hash_grid': 'L16 F2 T2^19 x2', 'mlp': '4x 64-wide (K=8)',
                   'occupancy': 'H128 C2 update every 16 steps', 'amp': amp, 'optimizer': 'Adam(fused)+EMA',
                   'l2_policy': 'no flush: per-step working set (~%d MB of samples/activations + 96 MB tables) exceeds the 126 MB L2'
                                % (samples_last * 700 // (1 << 20)),
                   'parallelism': ('dp%d (rays sharded; exchange fused into the optimizer kernel over NVLink peer memory: multimem.ld_reduce of '
                                   'the table-gradient shard, Adam on 1/N rows, multimem.st of the fp16 rows to every rank)' % world)
                                  if getattr(ts_p2p, 'value', False) else
                                  ('dp%d (rays sharded; table grads reduce-scattered, Adam on 1/N table shards, fp16 tables all-gathered; '
                                   'MLP grads all-reduced -- NCCL over NVLink)' % world)},
        'e2e': {'value': round(n_global * K / e2e_s, 1), 'unit': 'rays/s', 'h2d_bytes_per_step': int(host[0].numel() * 4),
                'd2h_bytes_per_step': 4 + 8, 'ms_per_step': round(e2e_s * 1e3 / K, 4)},
        'gpu_launches': int(launches),
        'clocks': clk,
        'roofline': roofline,
        'kernels': kern,
        'final_loss': lv,
        'cuda_mallocs_in_timed_region': int(seg1 - seg0),
    }
    if strong is not None:
        line['strong_scaling'] = strong
    if render_n is not None:
        line['render_full_frame'] = render_n
    if world > 1:
        # exchange time per step (CUDA events around every collective / peer-memory kernel on the stream it is issued on,
        # instrumented pass).  Peer-memory path (default): `barrier` (all ranks' gradients complete), `p2p_small` (MLP
        # gradients + found-inf flag), `p2p_adam_pair` (reduce-scatter + Adam + all-gather of the tables in ONE kernel,
        # so this entry also contains the optimizer arithmetic), `barrier_tail` (on a side stream under the next step's
        # ray marching; what the compute stream still waits for it is `wait_gather`).  NCCL path: reduce_scatter /
        # all_reduce_small / all_gather.  overlap_frac = hidden / total.
        total = sum(v for k, v in comm.items() if k != 'wait_gather')
        hidden = max(0.0, comm.get('all_gather', 0.0) + comm.get('barrier_tail', 0.0) - comm.get('wait_gather', 0.0))
        line['comm_ms'] = {k: round(v, 4) for k, v in comm.items()}
        line['comm_ms']['total'] = round(total, 4)
        line['overlap_frac'] = round(hidden / total, 3) if total > 0 else None
        line['exchange'] = {'peer_memory': bool(ts_p2p.value), 'nvls_multicast': bool(ts_p2p.multicast)}
    if world == 1 and not args.skip_cpu_baseline:
        line['cpu_baseline'] = cpu_baseline(args.cpu_sample_rays)
    if world == 1 and not args.skip_extras:
        try:
            line['extras'] = extras(device, peaks)
        except Exception as e:          # secondary evidence only; never fails the headline line
            line['extras'] = {'unavailable': '%s: %s' % (type(e).__name__, str(e)[:200])}
    if world == 1 and not args.skip_ref_ext:
        try:
            from bench_ref_ext import time_reference_ext
            line['reference_cuda_ext'] = time_reference_ext(args.rays, 16, device)
            from bench_ref_ext import time_reference_ext_render
            line['reference_cuda_ext']['render_full_frame'] = time_reference_ext_render(1008, 756, device)
        except Exception as e:      # the rebuilt reference extensions are optional evidence
            line['reference_cuda_ext'] = {'unavailable': '%s: %s' % (type(e).__name__, str(e)[:200])}
    print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


# ------------------------------------------------------------------------------------------------ secondary measurements
def render_sharded(device, rank, world, w=1008, h=756, frames=3):
    """Config 3 at N GPUs: each rank renders a contiguous block of rows of the 1008 x 756 frame with the device-driven
    loop, the tiles are gathered on rank 0 (39.6 MB); ms/frame = max over ranks, wall clock around render + gather."""
    import torch
    import torch.distributed as dist
    from nerfstyle_b200 import model as M, parallel, raymarching, scenes
    torch.manual_seed(0)
    m = M.StyleTCNerf([-2., -2., -2.], [2., 2., 2.], class_dim=N_CLASSES).to(device)
    r = M.Renderer(m, 2.0, raymarch_channels=3 + N_CLASSES, density_scale=50.0).to(device)
    r.density_bitfield = raymarching.packbits(scenes.analytic_density_grid(2, 128, 2.0).to(device), 0.5)
    intr = scenes.scaled_intrinsics(w, h)
    poses = scenes.synthetic_poses(frames + 1, 1)
    lo, hi = parallel.shard_bounds(w * h, rank, world)
    idx = torch.arange(lo, hi, device=device)
    ts = []
    for f in range(frames + 1):
        o, d = scenes.generate_rays(poses[f], intr, device, idx)
        dist.barrier()
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        with torch.no_grad(), torch.autocast('cuda', dtype=torch.float16):
            img, depth, cls = r.render_test_graph(o, d)
        full = parallel.gather_rows(torch.cat([img, depth[:, None], cls], dim=1), w * h, rank, world)
        if rank == 0:
            float(full[:, :3].sum().item())
        dist.barrier()
        torch.cuda.synchronize()
        if f > 0:
            ts.append(time.perf_counter() - t0)
    t = torch.tensor([sum(ts) / len(ts)], device=device, dtype=torch.float64)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    sec = float(t.item())
    return {'w': w, 'h': h, 'n_gpus': world, 'ms_per_frame': round(sec * 1e3, 2), 'mrays_per_s': round(w * h / sec / 1e6, 2),
            'case': 'analytic occupancy, density_scale 50; row tiles sharded, image gathered on rank 0'}


def extras(device, peaks):
    """The other two numbers BASELINE.json's metric names, measured in the same run (N = 1 only, a few seconds):
    full-frame render Mrays/s (config 3, trained-like case) and the tensor-core kernels against the measured bf16 peak."""
    import torch
    from nerfstyle_b200 import _lib, model as M, nnfm, raymarching, scenes
    out = {}
    tf_peak = float(peaks.get('bf16_tflops', 1590.0))
    # ---- config 3: 1008 x 756 frame through march_rays / composite_rays (analytic occupancy, density_scale 50)
    torch.manual_seed(0)
    m = M.StyleTCNerf([-2., -2., -2.], [2., 2., 2.], class_dim=N_CLASSES).to(device)
    r = M.Renderer(m, 2.0, raymarch_channels=3 + N_CLASSES, density_scale=50.0).to(device)
    r.density_bitfield = raymarching.packbits(scenes.analytic_density_grid(2, 128, 2.0).to(device), 0.5)
    intr = scenes.scaled_intrinsics(1008, 756)
    poses = scenes.synthetic_poses(4, 1)
    idx = torch.arange(0, 1008 * 756, device=device)
    ms = []
    for f in range(4):
        o, d = scenes.generate_rays(poses[f], intr, device, idx)
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        with torch.no_grad(), torch.autocast('cuda', dtype=torch.float16):
            img, depth, cls = r.render_test_graph(o, d)       # device-driven loop (CUDA graph), nerfstyle_b200/model.py
        float(img.sum().item())              # D2H read of the frame's checksum
        if f > 0:
            ms.append((time.perf_counter() - t0) * 1e3)
    out['render_full_frame'] = {'w': 1008, 'h': 756, 'ms_per_frame': round(sum(ms) / len(ms), 2),
                                'mrays_per_s': round(1008 * 756 / (sum(ms) / len(ms)) / 1e3, 2),
                                'case': 'analytic occupancy, density_scale 50 (trained-like early termination)',
                                'steps_per_iteration': 8}
    ms8 = []
    for f in range(3):      # the same frames with 4 marching steps per iteration (more, smaller iterations; fewer samples wasted past termination)
        o, d = scenes.generate_rays(poses[f + 1], intr, device, idx)
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        with torch.no_grad(), torch.autocast('cuda', dtype=torch.float16):
            img, depth, cls = r.render_test_graph(o, d, steps_per_iteration=4)
        float(img.sum().item())
        if f > 0:
            ms8.append((time.perf_counter() - t0) * 1e3)
    out['render_full_frame']['ms_per_frame_4_steps_per_iteration'] = round(sum(ms8) / len(ms8), 2)
    del m, r
    # ---- config 4: matching GEMM (tcgen05) at room size
    N1, N2, K = 11844, 15876, 768
    g = torch.Generator().manual_seed(0)
    a = torch.nn.functional.normalize(torch.randn(N1, K, generator=g), dim=1).to(device).half()
    b = torch.nn.functional.normalize(torch.randn(N2, K, generator=g), dim=1).to(device).half()
    preds = torch.randint(0, 8, (N1,), generator=g).to(device)
    clusters = (torch.arange(N2) * 8 // N2).to(device)
    for _ in range(2):
        nnfm.nn_match(a, b, preds, clusters, list(range(8)))
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10):
        nnfm.nn_match(a, b, preds, clusters, list(range(8)))
    e1.record()
    torch.cuda.synchronize()
    t = e0.elapsed_time(e1) / 10
    tf = 2.0 * N1 * N2 * K / t / 1e9
    # the reference's own composition (loss.py:32-36,199-214: fp16 matmul -> per-class inf mask loop -> amin), same operands
    def torch_composition():
        dists = 1.0 - a @ b.T
        for i in range(8):
            dists[torch.logical_and(*torch.meshgrid(preds == i, clusters != i, indexing='ij'))] = float('inf')
        return torch.amin(dists, dim=1)
    torch_composition()
    e0.record()
    for _ in range(3):
        torch_composition()
    e1.record()
    torch.cuda.synchronize()
    t_ref = e0.elapsed_time(e1) / 3
    out['nnfm_matching'] = {'N1': N1, 'N2': N2, 'K': K, 'ms': round(t, 3), 'reference_torch_composition_ms': round(t_ref, 3),
                            'roofline': {'bound': 'tensor', 'achieved': round(tf, 1), 'peak': tf_peak, 'unit': 'TFLOP/s',
                                         'frac': round(tf / tf_peak, 4), 'kernel': 'k_nnfm_gemm_tc (tcgen05, incl. operand packing)'}}
    # ---- the fused MLPs (tcgen05): forward + backward of the density net on 4 Mi rows
    B = 1 << 22
    x = torch.randn(B, 32, device=device).half()
    dy = (torch.randn(B, 1, device=device) * 0.01).half()
    prm = (torch.randn(64 * 32 + 16 * 64, device=device) * 0.1).half()
    y = torch.empty(B, 1, device=device, dtype=torch.float16)
    dx = torch.empty_like(x)
    dp = torch.zeros(prm.numel(), device=device)
    st = torch.cuda.current_stream().cuda_stream
    lib = _lib.lib()

    def fb():
        lib.nrf_mlp_forward(x.data_ptr(), 1, prm.data_ptr(), B, 32, 1, 1, 64, 1, 0, y.data_ptr(), 1, st)
        lib.nrf_mlp_backward(x.data_ptr(), 1, prm.data_ptr(), dy.data_ptr(), 1, B, 32, 1, 1, 64, 1, 0, 128.0, dx.data_ptr(), 1,
                             dp.data_ptr(), st)
    for _ in range(3):
        fb()
    e0.record()
    for _ in range(10):
        fb()
    e1.record()
    torch.cuda.synchronize()
    t = e0.elapsed_time(e1) / 10
    flop = 3 * 2 * (64 * 32 + 64 * 1) * B                 # un-padded MACs x 2, forward + 2x backward
    io = B * (32 * 2 + 2 + 32 * 2 + 2 + 32 * 2)           # x, y, x again, dy, dx
    out['mlp_density_fwd_bwd'] = {'rows': B, 'ms': round(t, 3), 'tflops_unpadded': round(flop / t / 1e9, 1),
                                  'io_gbs': round(io / t / 1e6, 1), 'hbm_frac': round(io / t / 1e6 / float(peaks.get('hbm_gbs', 6650.0)), 4),
                                  'kernel': 'k_mlp_fwd_tc + k_mlp_bwd_tc (tcgen05); bound by row traffic, not by the tensor pipe'}
    return out


# ------------------------------------------------------------------------------------------------ CPU oracle arms
def _cpu_step_fn(n_rays):
    """One train step of the CPU oracle port on `n_rays` rays of the same workload (fwd + bwd + Adam)."""
    import torch
    import oracle
    from oracle import field
    from nerfstyle_b200 import scenes
    torch.set_num_threads(os.cpu_count())
    of = field.OracleField(bound=2.0, n_classes=N_CLASSES, half=False, seed=0, table_std=1e-4)
    opt = torch.optim.Adam(list(of.params.values()), lr=0.01, eps=1e-15)
    intr = dict(scenes.ROOM)
    poses = scenes.synthetic_poses(intr['n_train'], 0)
    gen = torch.Generator().manual_seed(69420)
    # occupancy like update_state at step 0: density of the random-init field vs its mean
    bits = None

    def occupancy():
        H = 128
        grid = torch.zeros(2, H ** 3)
        coords = torch.from_numpy(oracle.morton3D_invert(torch.arange(H ** 3, dtype=torch.int32).numpy()).astype('float32'))
        xyz = 2 * coords / (H - 1) - 1
        for cas in range(2):
            b = min(2 ** cas, 2.0)
            hg = b / H
            p = xyz * (b - hg) + (torch.rand(xyz.shape, generator=gen) * 2 - 1) * hg
            with torch.no_grad():
                sig = torch.cat([of.forward(p[i:i + 262144]).reshape(-1) for i in range(0, p.shape[0], 262144)])
            grid[cas] = sig
        return oracle.packbits(grid.numpy(), min(float(grid.clamp(min=0).mean()), 10.0))

    state = {'bits': bits, 'it': 0}

    def step():
        if state['bits'] is None:
            state['bits'] = occupancy()
        pose = poses[state['it'] % len(poses)]
        idx = scenes.frame_indices(intr, n_rays, gen)
        o, d = scenes.generate_rays(pose, intr, 'cpu', idx)
        rgb, seg = scenes.synthetic_target(idx, intr, N_CLASSES)
        out = field.render_train(of, o.numpy(), d.numpy(), state['bits'], 2, 128, 2.0)
        loss = field.train_step_loss(out, rgb, seg)
        opt.zero_grad()
        loss.backward()
        opt.step()
        state['it'] += 1
        return float(loss.detach()), int(out['counter'][0])
    return step


def cpu_baseline(n_rays, budget_s=12.0, max_steps=16):
    """Bounded sample: steps of `n_rays` rays of the same workload until ~`budget_s` seconds of CPU work are timed."""
    step = _cpu_step_fn(n_rays)
    step()                      # warm-up (includes the one-off occupancy sweep)
    t0 = time.perf_counter()
    k, ns = 0, 0
    while k < max_steps and (k == 0 or time.perf_counter() - t0 < budget_s):
        loss, ns = step()
        k += 1
    dt = time.perf_counter() - t0
    return {'value': round(n_rays * k / dt, 2), 'unit': 'rays/s', 'cores': os.cpu_count(), 'kind': 'port',
            'sample': '%d steps x %d rays of the same workload (%d samples in the last one), full train steps (fwd+bwd+Adam) of the '
                      'CPU oracle (OpenMP C kernels + torch-CPU MLPs), %.1f s of CPU work' % (k, n_rays, ns, dt)}


def run_reference(args):
    rank = int(os.environ.get('RANK', '0'))
    if rank != 0:
        return
    n = 512
    step = _cpu_step_fn(n)
    W = max(1, min(args.warmup, 2))
    for _ in range(W):
        step()
    K = max(1, min(args.steps, 128))      # exactly the requested steps (0.5-0.6 s each); capped so any K ends within minutes
    t0 = time.perf_counter()
    ns = 0
    for _ in range(K):
        loss, ns = step()
    dt = time.perf_counter() - t0
    v = round(n * K / dt, 2)
    world = int(os.environ.get('WORLD_SIZE', '1'))
    line = {'impl': 'reference', 'metric': 'train_rays_per_s', 'value': v, 'unit': 'rays/s', 'n_gpus': world, 'steps': K,
            'warmup': W, 'ms_per_step': round(dt * 1e3 / K, 2), 'higher_is_better': True, 'scaling': 'weak', 'vs_baseline': None,
            'dtype': 'f32', 'data': 'synthetic',
            'config': {'workload': WORKLOAD, 'rays_per_step_sample': n, 'samples_per_step_last': ns,
                       'note': 'the reference path is CUDA-only; this arm is the CPU oracle port of its kernels on all host '
                               'cores, each step a bounded %d-ray sample of the 8192-ray workload' % n},
            'cpu_baseline': {'value': v, 'unit': 'rays/s', 'cores': os.cpu_count(), 'kind': 'port',
                             'sample': '%d steps x %d rays (fwd+bwd+Adam)' % (K, n)},
            'e2e': {'value': v, 'unit': 'rays/s', 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0}}
    print(json.dumps(line))


if __name__ == '__main__':
    a = parse()
    if a.impl == 'reference':
        run_reference(a)
    else:
        run_ours(a)
