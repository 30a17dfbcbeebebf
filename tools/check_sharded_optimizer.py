"""torchrun check (N >= 2 GPUs): the sharded optimizer step (reduce-scatter + Adam on 1/N table shards + all-gather of the
fp16 copies) equals the replicated one (all-reduce + full Adam) on identical per-rank gradients -- including a step that
must be skipped because ONE rank saw an inf -- and an end-to-end training run gives the same losses.
    python -m torch.distributed.run --nproc-per-node 2 --master-addr 127.0.0.1 tools/check_sharded_optimizer.py
(Element-wise comparison of TRAINED hash tables is ill-conditioned -- Adam with eps = 1e-15 moves a row whose gradient
is rounding noise by +-lr, and the float-atomics order differs from run to run -- hence the synthetic gradients.)"""
import os
import sys

import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench as B  # noqa: E402
from nerfstyle_b200 import model as M  # noqa: E402
from nerfstyle_b200.optim import FusedAdamEMA  # noqa: E402
from nerfstyle_b200.trainer import TrainStep  # noqa: E402


def main():
    world, rank, local = int(os.environ['WORLD_SIZE']), int(os.environ['RANK']), int(os.environ['LOCAL_RANK'])
    torch.cuda.set_device(local)
    dev = torch.device('cuda', local)
    dist.init_process_group('nccl', device_id=dev)
    ok = True
    # ---- 1. optimizer equivalence on synthetic gradients
    # tensors 0 and 3 share a shape: the sharded optimizer pairs them (interleaved fp16 copies + interleaved gradient buffer)
    shapes = [(1 << 21, 2), (3072,), ((1 << 20) + 8 * world, 2), (1 << 21, 2)]
    torch.manual_seed(0)
    init = [torch.randn(s, device=dev) * 0.1 for s in shapes]
    opts = []
    for shard in (True, False):
        ps = [torch.nn.Parameter(t.clone()) for t in init]
        opts.append((ps, FusedAdamEMA(ps, lr=0.01, lr_decay_steps=50, ema_decay=0.95, init_scale=1024.0, growth_interval=3,
                                      world_size=world, rank=rank, shard_big=shard, pair_tables=shard)))
    assert any(sh is not None for sh in opts[0][1].shard) and all(sh is None for sh in opts[1][1].shard)
    assert opts[0][1].pair_idx == (0, 3) and opts[1][1].pair_idx is None
    if rank == 0:
        p2p = opts[0][1]._p2p
        print('peer-memory exchange:', 'on (NVLS multicast %s)' % ('yes' if p2p['grad_mc'] else 'no') if p2p else
              'off (%s)' % opts[0][1]._p2p_error)
    gen = torch.Generator(device=dev).manual_seed(100 + rank)                    # every rank has its own gradients
    for it in range(7):
        grads = [torch.randn(s, device=dev, generator=gen) for s in shapes]
        if it in (3, 4) and rank == world - 1:
            # one rank, inside ANOTHER rank's shard; iteration 3 goes through .grad (NCCL path), iteration 4 through the
            # interleaved buffer (peer-memory path when available)
            grads[0][12345, 1] = float('inf')
        for ps, opt in opts:
            scale = float(opt.scale.item())
            opt.zero_grad()
            for i, (p, g) in enumerate(zip(ps, grads)):
                if opt.pair_idx is not None and i in opt.pair_idx and it % 2 == 0:
                    # what the paired scatter kernel does: accumulate into the optimizer's interleaved buffer, .grad stays None
                    opt.grad_pair_buffer()[:, opt.pair_idx.index(i)] += g * scale
                else:
                    p.grad = g * scale
            opt.step()
    opts[0][1].gather_master()
    for i, (a, b) in enumerate(zip(opts[0][0], opts[1][0])):
        d = float((a.detach() - b.detach()).abs().max())
        tol = 0.0 if world == 2 else 1e-6          # a two-term float sum does not depend on the order
        ok &= d <= tol
        ha, hb = getattr(a, '_nrf_half_copy', None), getattr(b, '_nrf_half_copy', None)
        if ha is not None and hb is not None:
            ok &= float((ha.float() - hb.float()).abs().max()) <= (0.0 if world == 2 else 1e-3)
            ha = ha.contiguous()
            ref = ha.clone()
            dist.broadcast(ref, src=0)
            ok &= bool(torch.equal(ref, ha))                                      # identical gathered tables on every rank
        if rank == 0:
            print('tensor %d %s: max|sharded - replicated| = %.3g' % (i, tuple(a.shape), d))
    ok &= int(opts[0][1].good_steps.item()) == 5 and int(opts[1][1].good_steps.item()) == 5     # the two inf steps were skipped
    ok &= float(opts[0][1].scale.item()) == float(opts[1][1].scale.item())
    # ---- 2. end to end: same losses
    host, devb = B.make_batches(6, 2048, rank, world, dev)
    losses = []
    for shard, pair in ((True, True), (False, True), (False, False)):
        torch.manual_seed(0)
        torch.cuda.manual_seed_all(0)
        m = M.StyleTCNerf([-2., -2., -2.], [2., 2., 2.], class_dim=B.N_CLASSES).to(dev)
        r = M.Renderer(m, 2.0, raymarch_channels=3 + B.N_CLASSES).to(dev)
        ts = TrainStep(r, enable_amp=True, world_size=world, shard_optimizer=shard, pair_tables=pair)
        losses.append([float(ts.step(*B.unpack(devb[s]))) for s in range(6)])
    ok &= all(abs(a - b) <= 2e-3 * abs(b) and abs(c - b) <= 2e-3 * abs(b) for a, b, c in zip(*losses))
    if rank == 0:
        print('losses sharded + paired   :', ['%.5f' % v for v in losses[0]])
        print('losses replicated + paired:', ['%.5f' % v for v in losses[1]])
        print('losses replicated, unpaired:', ['%.5f' % v for v in losses[2]])
    t = torch.tensor([1 if ok else 0], device=dev)
    dist.all_reduce(t, op=dist.ReduceOp.MIN)
    if rank == 0:
        print('SHARDED OPTIMIZER CHECK', 'OK' if int(t) == 1 else 'FAILED')
    dist.destroy_process_group()
    sys.exit(0 if int(t) == 1 else 1)


if __name__ == '__main__':
    main()
