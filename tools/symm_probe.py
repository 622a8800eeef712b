#!/usr/bin/env python
"""Does torch's symmetric memory (peer-mapped buffers + signal pads) work on this box, and what does a barrier cost?
    torchrun --nproc-per-node R tools/symm_probe.py"""
import os
import time

import torch
import torch.distributed as dist
import torch.distributed._symmetric_memory as symm_mem

rank, world, local = int(os.environ['RANK']), int(os.environ['WORLD_SIZE']), int(os.environ['LOCAL_RANK'])
torch.cuda.set_device(local)
dev = torch.device('cuda', local)
dist.init_process_group('nccl', device_id=dev)
n = 1 << 20
t = symm_mem.empty(world * n, dtype=torch.float32, device=dev)
hdl = symm_mem.rendezvous(t, dist.group.WORLD)
t.zero_()
hdl.barrier(channel=0)
# every rank writes its slab into every peer's buffer
src = torch.full((n,), float(rank + 1), device=dev)
for p in range(world):
    hdl.get_buffer(p, (world * n,), torch.float32)[rank * n:(rank + 1) * n].copy_(src)
hdl.barrier(channel=0)
got = t.view(world, n)[:, 0].tolist()
ok = got == [float(r + 1) for r in range(world)]
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(200):
    hdl.barrier(channel=0)
e1.record()
torch.cuda.synchronize()
bar_us = e0.elapsed_time(e1) / 200 * 1e3
x = torch.zeros(1, device=dev)
e0.record()
for _ in range(200):
    dist.all_reduce(x)
e1.record()
torch.cuda.synchronize()
ar_us = e0.elapsed_time(e1) / 200 * 1e3
# peer write bandwidth: 16 MB to one peer
big = symm_mem.empty(8 << 20, dtype=torch.float32, device=dev)
h2 = symm_mem.rendezvous(big, dist.group.WORLD)
peer = h2.get_buffer((rank + 1) % world, (8 << 20,), torch.float32)
srcb = torch.ones(8 << 20, device=dev)
for _ in range(3):
    peer.copy_(srcb)
torch.cuda.synchronize()
e0.record()
for _ in range(20):
    peer.copy_(srcb)
e1.record()
torch.cuda.synchronize()
gbs = 20 * (32 << 20) / (e0.elapsed_time(e1) * 1e-3) / 1e9
print(f'rank {rank}: peer writes ok={ok} ptrs={len(hdl.buffer_ptrs)} '
      f'mc_ptr={hdl.multicast_ptr} barrier {bar_us:.1f} us, 4-byte all_reduce {ar_us:.1f} us, peer copy {gbs:.0f} GB/s', flush=True)
dist.barrier()
dist.destroy_process_group()
