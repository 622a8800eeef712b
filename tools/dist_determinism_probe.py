#!/usr/bin/env python
"""Is the sharded head run-to-run bit-identical, and is prefetch(labels) + forward_pair(x, y) bit-identical to
forward_pair(x, y, labels)?      torchrun --nproc-per-node R tools/dist_determinism_probe.py
Four fresh heads (plain, plain, prefetch, prefetch) run the same seeded steps; rank 0 prints the comparison."""
import json
import os
import sys

R_ = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [R_, os.path.join(R_, 'very-large-scale-face-recognition_b200')]
import torch
import torch.distributed as dist
import torch.nn.functional as F
from ffc_b200.dist import ShardedFFCHead

rank, world, local = int(os.environ['RANK']), int(os.environ['WORLD_SIZE']), int(os.environ['LOCAL_RANK'])
torch.cuda.set_device(local)
dev = torch.device('cuda', local)
dist.init_process_group('nccl', device_id=dev)
D, Q, B, N, steps = 512, int(os.environ.get('PROBE_Q', 262144)), 1024, int(os.environ.get('PROBE_N', 300000)), 12
g = torch.Generator().manual_seed(77 + rank)
batches = []
for s in range(steps):
    h = B // 2
    ids = torch.randint(0, N, (h,), generator=torch.Generator().manual_seed(1000 + s))[: h]
    xl = torch.cat([ids, torch.randint(0, N, (B - h,), generator=g)]).pin_memory()
    yl = torch.cat([ids, torch.randint(0, N, (B - h,), generator=g)]).pin_memory()
    batches.append((F.normalize(torch.randn(B, D, generator=g)).to(dev), F.normalize(torch.randn(B, D, generator=g)).to(dev), xl, yl))


def run(prefetch):
    torch.manual_seed(5)
    head = ShardedFFCHead(D, Q, 32.0, 'Arc', 0.5, max_batch=B, device=dev)
    head.prefill_identity(min(N, Q))
    out = []
    for s, (x, y, xl, yl) in enumerate(batches):
        loss, dx, dy = head.forward_pair(x, y, xl, yl)
        if prefetch and s + 1 < steps:
            head.prefetch(batches[s + 1][2], batches[s + 1][3])
        out.append((loss.clone(), dx.clone(), dy.clone()))
    torch.cuda.synchronize()
    lru = head.backend.lru.state_dict()
    q = head.backend.queue.clone()
    return out, lru, q, head.prefetch_hits


runs = [run(False), run(False), run(True), run(True)]


def same(a, b):
    ok = a[1] == b[1] and torch.equal(a[2], b[2])
    first = None
    for s, (u, v) in enumerate(zip(a[0], b[0])):
        if not all(torch.equal(p, q) for p, q in zip(u, v)):
            ok = False
            if first is None:
                first = (s, float(u[0]), float(v[0]), float((u[1] - v[1]).abs().max()))
    return ok, first


res = dict(world=world, plain_vs_plain=same(runs[0], runs[1]), prefetch_vs_prefetch=same(runs[2], runs[3]), plain_vs_prefetch=same(runs[0], runs[2]),
           prefetch_hits=runs[2][3], lru_equal=runs[0][1] == runs[2][1])
flags = torch.tensor([int(res[k][0]) for k in ('plain_vs_plain', 'prefetch_vs_prefetch', 'plain_vs_prefetch')], device=dev)
dist.all_reduce(flags, op=dist.ReduceOp.MIN)
if rank == 0:
    res['all_ranks'] = flags.tolist()
    print(json.dumps(res))
dist.destroy_process_group()
