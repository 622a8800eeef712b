"""Generate golden vectors by running the UNMODIFIED reference (/root/reference) on CPU.

Run in the build container only (the reference does not travel to the GPU box):
    python tests/golden/make_golden.py
Writes tests/golden/lru_traces.json and tests/golden/ffc_<case>.npz (committed).
The shims are described in oracle/ref_shim.py; no reference source is copied.
"""
import json
import os
import random
import sys

import numpy as np
import torch
import torch.nn.functional as F

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import ref_shim  # noqa: E402

OUT = os.path.dirname(os.path.abspath(__file__))


def lru_traces():
    _, lru_mod = ref_shim.load()
    rng = random.Random(1234)
    cases = []
    # (capacity, key universe, number of ops)
    for cap, universe, n_ops in [(1, 4, 60), (2, 5, 80), (4, 9, 200), (7, 10, 300), (16, 64, 600), (64, 96, 1500),
                                 (64, 4096, 800)]:
        lru = lru_mod.LRU(cap)
        ops = []
        pending = 0
        for _ in range(n_ops):
            r = rng.random()
            key = rng.randrange(universe)
            if r < 0.45 and pending == 0:
                ops.append(['get', key, lru.get(key)])
            elif r < 0.70:
                ops.append(['try_get', key, lru.try_get(key)])
                pending += 1
            elif r < 0.80 and pending > 0:
                n = rng.randrange(1, pending + 1) if rng.random() < 0.5 else pending
                lru.rollback_steps(n)
                pending -= n
                ops.append(['rollback_steps', n, None])
            elif r < 0.90:
                ops.append(['view', key, lru.view(key)])
            elif r < 0.95:
                ops.append(['contains', key, key in lru])
            else:
                ops.append(['state', None, [list(kv) for kv in lru.state_dict()], lru.cur_idx])
        if pending:
            lru.rollback_steps(pending)
            ops.append(['rollback_steps', pending, None])
        ops.append(['state', None, [list(kv) for kv in lru.state_dict()], lru.cur_idx])
        cases.append(dict(capacity=cap, ops=ops))
    # restore round trip
    lru = lru_mod.LRU(8)
    for k in (5, 9, 2, 5, 7, 1, 9):
        lru.get(k)
    sd = lru.state_dict()
    lru2 = lru_mod.LRU(8)
    lru2.restore(sd)
    after = [lru2.get(k) for k in (3, 4, 6, 8, 10, 5)]
    cases.append(dict(capacity=8, restore=[list(kv) for kv in sd], gets=[3, 4, 6, 8, 10, 5], slots=after,
                      final=[list(kv) for kv in lru2.state_dict()]))
    with open(os.path.join(OUT, 'lru_traces.json'), 'w') as f:
        json.dump(cases, f, separators=(',', ':'))
    print('lru_traces.json', sum(len(c.get('ops', [])) for c in cases), 'ops')


def centers(n_ids, D, seed):
    g = torch.Generator().manual_seed(seed)
    return F.normalize(torch.randn(n_ids, D, generator=g))


def make_batch(gen, cen, B, n_ids, noise):
    """main.py:53-60 composition: id half (same ids in x and y) + two independent instance halves."""
    h = B // 2
    ids = torch.randperm(n_ids, generator=gen)[:h]
    ins1 = torch.randint(0, n_ids, (B - h,), generator=gen)
    ins2 = torch.randint(0, n_ids, (B - h,), generator=gen)
    xl = torch.cat([ids, ins1])
    yl = torch.cat([ids, ins2])
    D = cen.shape[1]
    x = F.normalize(cen[xl] + noise * torch.randn(B, D, generator=gen))
    y = F.normalize(cen[yl] + noise * torch.randn(B, D, generator=gen))
    return x, y, xl.tolist(), yl.tolist()


def ffc_case(name, D, Q, B, n_ids, steps, loss_type, margin, scale, noise, seed):
    torch.manual_seed(seed)
    m = ref_shim.make_ffc(D, Q, scale, loss_type, margin)
    q0 = m.queue.detach().clone()
    gen = torch.Generator().manual_seed(seed + 1)
    cen = centers(n_ids, D, seed + 2)
    rec = dict(D=D, Q=Q, B=B, n_ids=n_ids, steps=steps, loss_type=loss_type, margin=margin, scale=scale,
               queue0=q0.numpy())
    for s in range(steps):
        x, y, xl, yl = make_batch(gen, cen, B, n_ids, noise)
        loss, dx, dy, tr = ref_shim.forward_backward(m, x, y, xl, yl)
        rec[f'x{s}'], rec[f'y{s}'] = x.numpy(), y.numpy()
        rec[f'xl{s}'], rec[f'yl{s}'] = np.array(xl), np.array(yl)
        rec[f'loss{s}'] = np.float64(loss)
        rec[f'dx{s}'], rec[f'dy{s}'] = dx.numpy(), dy.numpy()
        for pi, pn in enumerate(('rb', 'cm')):
            for k in ('rows', 'cols', 'labels', 'ones'):
                rec[f'{pn}_{k}{s}'] = np.array(tr[pi][k], dtype=np.int64)
        rec[f'lru{s}'] = np.array(m.lru.state_dict(), dtype=np.int64).reshape(-1, 2)
        rec[f'qpos{s}'] = np.array([m.queue_position_dict[i] for i in range(Q)], dtype=np.int64)
    rec['queue_final'] = m.queue.detach().numpy()
    np.savez_compressed(os.path.join(OUT, f'ffc_{name}.npz'), **rec)
    print(name, 'losses', [float(rec[f'loss{s}']) for s in range(steps)])


if __name__ == '__main__':
    # serial index_put => deterministic "last duplicate wins" (probed)
    torch.set_num_threads(1)
    lru_traces()
    # small queue with evictions + in-batch duplicates + ids colliding between halves
    ffc_case('am_evict', D=16, Q=48, B=16, n_ids=80, steps=6, loss_type='AM', margin=0.4, scale=32.0, noise=0.3, seed=11)
    ffc_case('arc_evict', D=16, Q=48, B=16, n_ids=80, steps=6, loss_type='Arc', margin=0.5, scale=32.0, noise=0.3, seed=12)
    ffc_case('sv_evict', D=16, Q=48, B=16, n_ids=80, steps=6, loss_type='SV', margin=0.4, scale=32.0, noise=0.3, seed=13)
    # capacity smaller than the batch: slot reuse inside one batch
    ffc_case('am_tiny', D=8, Q=8, B=16, n_ids=40, steps=4, loss_type='AM', margin=0.4, scale=16.0, noise=0.3, seed=14)
    # queue == identity count (steady state all hits), ragged odd batch
    ffc_case('arc_full', D=32, Q=64, B=13, n_ids=64, steps=8, loss_type='Arc', margin=0.5, scale=32.0, noise=0.25, seed=15)
    # shapes the tensor-core path tiles: D=128, Q not a multiple of 128
    ffc_case('arc_d128', D=128, Q=1000, B=96, n_ids=2500, steps=5, loss_type='Arc', margin=0.5, scale=32.0, noise=0.6, seed=16)
    ffc_case('am_d128', D=128, Q=640, B=64, n_ids=700, steps=6, loss_type='AM', margin=0.4, scale=32.0, noise=0.6, seed=17)
