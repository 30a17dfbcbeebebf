"""Probe (torchrun, N >= 2): does torch.distributed._symmetric_memory give peer pointers usable from our own kernels?"""
import os
import sys
import torch
import torch.distributed as dist

rank, world, local = int(os.environ['RANK']), int(os.environ['WORLD_SIZE']), int(os.environ['LOCAL_RANK'])
torch.cuda.set_device(local)
dev = torch.device('cuda', local)
dist.init_process_group('nccl', device_id=dev)
import torch.distributed._symmetric_memory as symm_mem
print(rank, 'symm_mem attrs', [a for a in dir(symm_mem) if not a.startswith('_')][:40], flush=True)
t = symm_mem.empty(1 << 20, dtype=torch.float32, device=dev)
t.fill_(float(rank + 1))
hdl = symm_mem.rendezvous(t, dist.group.WORLD)
print(rank, 'handle', type(hdl), [a for a in dir(hdl) if not a.startswith('_')], flush=True)
ptrs = list(hdl.buffer_ptrs)
print(rank, 'buffer_ptrs', [hex(p) for p in ptrs], 'mine', hex(t.data_ptr()), flush=True)
hdl.barrier()
peer = hdl.get_buffer((rank + 1) % world, (1 << 20,), torch.float32)
print(rank, 'peer value', float(peer[0]), float(peer[-1]), flush=True)
# raw-pointer access from a plain tensor view is what our kernels will do; time a P2P read of the whole peer buffer
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
big = symm_mem.empty(25 << 20, dtype=torch.float32, device=dev)      # 100 MB
hb = symm_mem.rendezvous(big, dist.group.WORLD)
big.fill_(1.0)
hb.barrier()
pb = hb.get_buffer((rank + 1) % world, (25 << 20,), torch.float32)
out = torch.empty_like(big)
for _ in range(3):
    out.copy_(pb)
e0.record()
for _ in range(10):
    out.copy_(pb)
e1.record()
torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / 10
print(rank, 'P2P read 100 MB: %.3f ms = %.0f GB/s' % (ms, 100 * 1.048576 / ms), flush=True)
e0.record()
for _ in range(10):
    hb.barrier()
e1.record()
torch.cuda.synchronize()
print(rank, 'symm barrier: %.1f us' % (e0.elapsed_time(e1) * 100), flush=True)
dist.barrier()
dist.destroy_process_group()
