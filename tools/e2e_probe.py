#!/usr/bin/env python
"""Where does a step go when labels come from the host?  python tools/e2e_probe.py [c4|c3|c2]
Times forward_pair with device labels / pinned host labels, synchronising per step (host wall clock and CUDA events)."""
import os
import sys
import time

R = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [R, os.path.join(R, 'very-large-scale-face-recognition_b200')]
import importlib.util

import torch

spec = importlib.util.spec_from_file_location('bench_mod', os.path.join(R, 'bench.py'))
b = importlib.util.module_from_spec(spec)
spec.loader.exec_module(b)
import ffc_b200

w = b.WORKLOADS[sys.argv[1] if len(sys.argv) > 1 else 'c4']
dev = torch.device('cuda')
B, N, Q, D = w['B'], w['N'], w['Q'], w['D']
head = ffc_b200.FFCHead(D, Q, w['scale'], w['loss_type'], w['margin'], precision='bf16', max_batch=B, device=dev)
head._ensure()
head.lru.restore_arrays(torch.arange(Q, dtype=torch.int64), torch.arange(Q, dtype=torch.int32))
host = b.make_batches(w, 24, seed=1234)
for mode in ('device labels', 'pinned host labels', 'device labels', 'pinned host labels'):
    bs = [(x.to(dev), y.to(dev), xl.to(dev) if mode.startswith('device') else xl.pin_memory(), yl.to(dev) if mode.startswith('device') else yl.pin_memory())
          for x, y, xl, yl in host]
    torch.cuda.synchronize()
    ts, es = [], []
    for i, (x, y, xl, yl) in enumerate(bs):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        t0 = time.perf_counter()
        e0.record()
        head.forward_pair(x, y, y, x, xl, yl)
        t1 = time.perf_counter()
        e1.record()
        torch.cuda.synchronize()
        t2 = time.perf_counter()
        if i >= 4:
            ts.append(((t1 - t0) * 1e3, (t2 - t0) * 1e3))
            es.append(e0.elapsed_time(e1))
    print(f'{mode:22s} per-step sync: host enqueue {sum(t[0] for t in ts) / len(ts):.3f} ms, host total {sum(t[1] for t in ts) / len(ts):.3f} ms, '
          f'device {sum(es) / len(es):.3f} ms')
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for x, y, xl, yl in bs:
        head.forward_pair(x, y, y, x, xl, yl)
    torch.cuda.synchronize()
    print(f'{mode:22s} free-running: {(time.perf_counter() - t0) * 1e3 / len(bs):.3f} ms/step')

# the regime of round 1's e2e leg: the same batches again, newest first (their x-side identities were committed, the y-side instance
# identities were not: the rollback pass mixes hits of recent entries with misses)
bs = [(x.to(dev), y.to(dev), xl.pin_memory(), yl.pin_memory()) for x, y, xl, yl in host][::-1]
torch.cuda.synchronize()
for rep in range(2):
    ts = []
    for x, y, xl, yl in bs:
        t0 = time.perf_counter()
        head.forward_pair(x, y, y, x, xl, yl)
        torch.cuda.synchronize()
        ts.append((time.perf_counter() - t0) * 1e3)
    print(f're-fed batches (newest first), pass {rep}: mean {sum(ts) / len(ts):.3f} ms/step, max {max(ts):.3f}, first five {[round(t, 2) for t in ts[:5]]}')
