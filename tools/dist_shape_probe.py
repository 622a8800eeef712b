#!/usr/bin/env python
"""Per-phase device time of ONE rank of the R-way sharded head at its real shapes (R*B rows, Q/R columns) on a single GPU:
the collectives are left out (records are replicated locally), everything else is the product path of ffc_b200/dist.py.
    python tools/dist_shape_probe.py [R] [B] [Q] [D]"""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, 'very-large-scale-face-recognition_b200')]
from ffc_b200.dist import CudaShardBackend   # noqa: E402
from ffc_b200.ffc import hard_neg_k          # noqa: E402

R, B, Q, D = (int(a) for a in (sys.argv[1:5] + ['8', '1024', str(1 << 20), '512'][len(sys.argv) - 1:]))
OUTLIERS = os.environ.get('PROBE_OUTLIERS') == '1'     # 1: probe rows owned by other ranks count as unknown (label -1): top-k stress
dev = torch.device('cuda', 0)
n, Ql = R * B, Q // R
be = CudaShardBackend(D, Ql, Q, 0, n, 32.0, 'Arc', 0.5, hard_neg_k(Q), 'bf16', dev)
be.lru.restore_arrays(torch.arange(0, Q, R, dtype=torch.int64)[:Ql], torch.arange(Ql, dtype=torch.int32))
g = torch.Generator().manual_seed(0)
rec = be.new_records(n, R)
phases = {}


def timed(name, fn):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    out = fn()
    e1.record()
    phases.setdefault(name, []).append((e0, e1))
    return out


for it in range(12):
    keys = torch.randint(0, Q, (n,), generator=g).to(dev)
    probe = torch.randint(0, Q, (n,), generator=g).to(dev)
    p = torch.nn.functional.normalize(torch.randn(n, D, generator=g)).to(dev)
    gal = torch.nn.functional.normalize(torch.randn(n, D, generator=g)).to(dev)
    for commit in (False, True):
        be.use_set(1 if commit else 0)
        keys_c, order, n_mine = timed('route', lambda: be.route(keys, R, 0))
        timed('lru_assign', lambda: be.assign(keys_c, n_mine, journal=not commit))
        loc = timed('view', lambda: be.view(probe))
        other = torch.full_like(loc, -1) if OUTLIERS else (torch.remainder(probe, R) * Ql + probe // R).to(loc.dtype)   # resident on a peer
        label = torch.where((loc >= 0) & (torch.remainder(probe, R) == 0), loc, other).to(torch.int32)
        if not commit:
            timed('lru_undo', be.undo_bookkeeping)
        timed('scatter', lambda: be.scatter(gal, order, save_undo=not commit))
        timed('sweep_record', lambda: be.sweep_record(p, label, rec))
        rec['all'].copy_(rec['own'].unsqueeze(0).expand(R, -1, -1))
        timed('finalize_gathered', lambda: be.finalize_gathered(p, label, rec, R))
        if not commit:
            timed('restore', be.restore_queue)
        timed('end_pass', be.end_pass)
torch.cuda.synchronize()
print(f'R={R} B={B} Q={Q} D={D}: rows {n}, local columns {Ql}; us per call (median of the last 8 steps)')
for name, evs in phases.items():
    t = sorted(a.elapsed_time(b) * 1e3 for a, b in evs[len(evs) // 3:])
    print(f'  {name:18s} {t[len(t) // 2]:9.1f}')
