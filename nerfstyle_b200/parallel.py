"""Data parallelism over the rays of a step: one process per GPU, rays sharded, replicated
tables / MLPs / occupancy, one summed exchange of the parameter gradients per step (NCCL over NVLink on the GPU
box; the same code runs on gloo for the CPU tests).  Render tiles shard the same way with no exchange but the
final gather of the image rows.
"""
import torch
import torch.distributed as dist


def shard_bounds(n, rank, world):
    """Contiguous shard [lo, hi) of n units for `rank` of `world`; sizes differ by at most one, all units covered once."""
    base, rem = divmod(n, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def shard_rays(rays_o, rays_d, rank, world, *extra):
    lo, hi = shard_bounds(rays_o.shape[0], rank, world)
    return (rays_o[lo:hi], rays_d[lo:hi]) + tuple(e[lo:hi] for e in extra)


def local_loss_weight(n_local, n_global):
    """A rank's mean loss over its shard must be weighted by n_local / n_global so that the SUM over ranks of the
    gradients equals the gradient of the mean loss over the whole batch."""
    return float(n_local) / float(n_global)


def allreduce_grads(params, world, bucket_small_below=1 << 20):
    """Sum the gradients of `params` over all ranks in place.  Large tensors (the two 50 MB hash-table gradients) go
    on their own; the small MLP gradients are flattened into one message."""
    if world <= 1:
        return 0
    nbytes = 0
    big = [p.grad for p in params if p.grad is not None and p.grad.numel() >= bucket_small_below]
    small = [p.grad for p in params if p.grad is not None and p.grad.numel() < bucket_small_below]
    for g in big:
        dist.all_reduce(g, op=dist.ReduceOp.SUM)
        nbytes += g.numel() * g.element_size()
    if small:
        flat = torch.cat([g.reshape(-1) for g in small])
        dist.all_reduce(flat, op=dist.ReduceOp.SUM)
        nbytes += flat.numel() * flat.element_size()
        o = 0
        for g in small:
            g.copy_(flat[o:o + g.numel()].view_as(g))
            o += g.numel()
    return nbytes


def _is_nccl():
    return dist.get_backend() == 'nccl'


def allreduce_tensors(tensors, world, bucket_small_below=1 << 20):
    """Sum a list of tensors over all ranks in place (big ones on their own, small ones as one flattened message)."""
    if world <= 1 or not tensors:
        return
    small = [t for t in tensors if t.numel() < bucket_small_below]
    for t in tensors:
        if t.numel() >= bucket_small_below:
            dist.all_reduce(t, op=dist.ReduceOp.SUM)
    if small:
        flat = torch.cat([t.reshape(-1) for t in small])
        dist.all_reduce(flat, op=dist.ReduceOp.SUM)
        o = 0
        for t in small:
            t.copy_(flat[o:o + t.numel()].view_as(t))
            o += t.numel()


def reduce_scatter_sum(full, shard_out, world, rank):
    """shard_out <- this rank's 1/world slice of the sum of `full` over all ranks (full.numel() % world == 0).
    NCCL: one reduce-scatter; gloo (CPU tests) has none: all-reduce + slice."""
    per = full.numel() // world
    if _is_nccl():
        dist.reduce_scatter_tensor(shard_out, full, op=dist.ReduceOp.SUM)
    else:
        dist.all_reduce(full, op=dist.ReduceOp.SUM)
        shard_out.copy_(full[rank * per:(rank + 1) * per])


def all_gather_shards(full, shard, world):
    """full <- concatenation of every rank's `shard` (which may be the matching slice of `full` itself: in place)."""
    if _is_nccl():
        dist.all_gather_into_tensor(full, shard)
    else:
        parts = [torch.empty_like(shard) for _ in range(world)]
        dist.all_gather(parts, shard.clone())
        full.copy_(torch.cat(parts))


def allreduce_max_int(t, world):
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)


def gather_rows(local, n_total, rank, world, dst=0):
    """Gather row shards (render tiles) on `dst`; returns the full [n_total, ...] tensor there, None elsewhere."""
    if world <= 1:
        return local
    sizes = [shard_bounds(n_total, r, world) for r in range(world)]
    if rank == dst:
        out = local.new_empty((n_total,) + tuple(local.shape[1:]))
        out[sizes[dst][0]:sizes[dst][1]] = local
        for r in range(world):
            if r != dst:
                buf = local.new_empty((sizes[r][1] - sizes[r][0],) + tuple(local.shape[1:]))
                dist.recv(buf, src=r)
                out[sizes[r][0]:sizes[r][1]] = buf
        return out
    dist.send(local.contiguous(), dst=dst)
    return None
