"""Data-parallel host logic on CPU: world_size-2 gloo.  The sum over ranks of the shard gradients (loss weighted by
n_local / n_global) equals the whole-batch gradient, computed with the CPU oracle pipeline; render tiles gather back in
order.  (The CUDA ops need a GPU; what is tested here is the sharding / exchange code the GPU path uses.)"""
import os
import socket

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _free_port():
    s = socket.socket()
    s.bind(('127.0.0.1', 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _scene():
    import oracle
    from nerfstyle_b200 import scenes
    H = 32
    grid = scenes.analytic_density_grid(2, H, 2.0)
    bits = oracle.packbits(grid.numpy(), 0.5)
    o, d = scenes.random_rays(49, 3)       # odd count: uneven shards
    g = torch.Generator().manual_seed(1)
    return H, bits, o, d, torch.rand(49, 3, generator=g), torch.randint(0, 8, (49,), generator=g)


def _grads(o, d, tgt, cls, H, bits, weight):
    from oracle import field
    of = field.OracleField(bound=2.0, n_classes=8, half=False, seed=0, table_std=0.3, log2_hashmap_size=12, num_levels=8)
    out = field.render_train(of, o.numpy(), d.numpy(), bits, 2, H, 2.0, max_steps=128)
    loss = field.train_step_loss(out, tgt, cls) * weight
    loss.backward()
    return of, out


def _worker(rank, world, port, ret):
    os.environ['MASTER_ADDR'] = '127.0.0.1'
    os.environ['MASTER_PORT'] = str(port)
    dist.init_process_group('gloo', rank=rank, world_size=world)
    torch.set_num_threads(2)
    from nerfstyle_b200 import parallel
    H, bits, o, d, tgt, cls = _scene()
    so, sd, st, sc = parallel.shard_rays(o, d, rank, world, tgt, cls)
    of, out = _grads(so, sd, st, sc, H, bits, parallel.local_loss_weight(so.shape[0], o.shape[0]))

    class P:            # allreduce_grads only touches .grad
        def __init__(self, g):
            self.grad = g
    names = sorted(of.params)
    params = [P(of.params[n].grad) for n in names]
    nbytes = parallel.allreduce_grads(params, world, bucket_small_below=1 << 14)
    img = parallel.gather_rows(out['rgb'].detach(), o.shape[0], rank, world)
    if rank == 0:
        ret['grads'] = {n: p.grad.clone() for n, p in zip(names, params)}
        ret['img'] = img
        ret['nbytes'] = nbytes
    dist.barrier()
    dist.destroy_process_group()


def test_shard_bounds():
    from nerfstyle_b200.parallel import shard_bounds
    for n in (0, 1, 7, 8192, 762048):
        for w in (1, 2, 3, 8):
            b = [shard_bounds(n, r, w) for r in range(w)]
            assert b[0][0] == 0 and b[-1][1] == n
            assert all(b[i][1] == b[i + 1][0] for i in range(w - 1))
            sizes = [hi - lo for lo, hi in b]
            assert max(sizes) - min(sizes) <= 1


def test_dp2_gradients_equal_whole_batch():
    world = 2
    mgr = mp.Manager()
    ret = mgr.dict()
    mp.spawn(_worker, args=(world, _free_port(), ret), nprocs=world, join=True)
    H, bits, o, d, tgt, cls = _scene()
    of, out = _grads(o, d, tgt, cls, H, bits, 1.0)
    assert ret['nbytes'] > 0
    for n in sorted(of.params):
        a, b = ret['grads'][n].numpy(), of.params[n].grad.numpy()
        assert np.abs(b).max() > 0, n
        np.testing.assert_allclose(a, b, rtol=2e-4, atol=2e-6 * np.abs(b).max(), err_msg=n)
    np.testing.assert_allclose(ret['img'].numpy(), out['rgb'].detach().numpy(), rtol=1e-6, atol=1e-7)


def _shard_worker(rank, world, port, ret):
    os.environ['MASTER_ADDR'] = '127.0.0.1'
    os.environ['MASTER_PORT'] = str(port)
    dist.init_process_group('gloo', rank=rank, world_size=world)
    from nerfstyle_b200 import parallel
    n = 24
    full = torch.arange(n, dtype=torch.float32) * (rank + 1)              # rank r holds (r+1) * [0..n)
    shard = torch.empty(n // world)
    parallel.reduce_scatter_sum(full.clone(), shard, world, rank)          # sum over ranks = 3 * [0..n) for world 2
    table = torch.zeros(n)
    per = n // world
    table[rank * per:(rank + 1) * per] = shard                             # "update" only the own shard ...
    parallel.all_gather_shards(table, table[rank * per:(rank + 1) * per], world)   # ... then gather in place
    flag = torch.tensor([1 if rank == 1 else 0], dtype=torch.int32)
    parallel.allreduce_max_int(flag, world)
    small = [torch.full((3,), float(rank + 1)), torch.full((2, 2), 10.0 * (rank + 1))]
    parallel.allreduce_tensors(small, world, bucket_small_below=1 << 10)
    if rank == 0:
        ret['table'] = table
        ret['flag'] = int(flag)
        ret['small'] = [t.clone() for t in small]
    dist.barrier()
    dist.destroy_process_group()


def test_sharded_exchange_primitives():
    """The collectives behind the sharded optimizer step (reduce-scatter of the gradient, in-place all-gather of the
    updated shards, MAX of the found-inf flag, flattened all-reduce of the small tensors) on world_size-2 gloo."""
    world = 2
    mgr = mp.Manager()
    ret = mgr.dict()
    mp.spawn(_shard_worker, args=(world, _free_port(), ret), nprocs=world, join=True)
    torch.testing.assert_close(ret['table'], torch.arange(24, dtype=torch.float32) * 3)
    assert ret['flag'] == 1
    torch.testing.assert_close(ret['small'][0], torch.full((3,), 3.0))
    torch.testing.assert_close(ret['small'][1], torch.full((2, 2), 30.0))


def _pair_worker(rank, world, port, ret):
    os.environ['MASTER_ADDR'] = '127.0.0.1'
    os.environ['MASTER_PORT'] = str(port)
    dist.init_process_group('gloo', rank=rank, world_size=world)
    from nerfstyle_b200 import parallel
    T = 12                                                  # rows per table
    # interleaved gradient buffer [row][table][2]: table e of rank r holds (r+1) * (100 e + 2 row + feature)
    row = torch.arange(T, dtype=torch.float32)[:, None, None]
    e = torch.arange(2, dtype=torch.float32)[None, :, None]
    f = torch.arange(2, dtype=torch.float32)[None, None, :]
    gp = (100 * e + 2 * row + f) * (rank + 1)
    per = 2 * T // world                                    # parameter elements per rank and table (FusedAdamEMA.shard)
    lo = rank * per
    shard = torch.empty(gp.numel() // world)
    parallel.reduce_scatter_sum(gp.reshape(-1).clone(), shard, world, rank)
    # the shard of the interleaved buffer is exactly this rank's ROW range of both tables: rows [lo/2, (lo+per)/2)
    want = (gp * 3 / (rank + 1))[lo // 2:(lo + per) // 2]   # sum over ranks 1 + 2 = 3
    ok = bool(torch.equal(shard.view(-1, 2, 2), want))
    # "update" the fp16 pair buffer on the own rows only, then gather: elements [2 lo, 2 (lo + per)) of the flat buffer
    half_pair = torch.zeros(T, 2, 2, dtype=torch.float16)
    half_pair[lo // 2:(lo + per) // 2] = (shard.view(-1, 2, 2) / 3).half()
    flat = half_pair.view(-1)
    parallel.all_gather_shards(flat, flat[2 * lo:2 * (lo + per)], world)
    if rank == 0:
        ret['ok'] = ok
        ret['half_pair'] = half_pair.clone()
    dist.barrier()
    dist.destroy_process_group()


def test_paired_table_exchange_index_arithmetic():
    """The data-parallel step with paired tables: ONE reduce-scatter of the interleaved gradient buffer hands every rank the
    same ROW range of both tables, ONE in-place all-gather of the interleaved fp16 buffer rebuilds both (optim.py)."""
    world = 2
    mgr = mp.Manager()
    ret = mgr.dict()
    mp.spawn(_pair_worker, args=(world, _free_port(), ret), nprocs=world, join=True)
    assert ret['ok']
    row = torch.arange(12, dtype=torch.float32)[:, None, None]
    e = torch.arange(2, dtype=torch.float32)[None, :, None]
    f = torch.arange(2, dtype=torch.float32)[None, None, :]
    torch.testing.assert_close(ret['half_pair'].float(), 100 * e + 2 * row + f)
