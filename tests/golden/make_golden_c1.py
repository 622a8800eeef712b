"""C1-shaped fixtures (BASELINE.json configs[0]: batch 64, 10k identities, queue 4096; feat_dim 128 = MobileFaceNet's native width,
mobilefacenet_def.py:78, and 512 = main.py:163's default) from the UNMODIFIED reference head on CPU, embeddings in.

    python tests/golden/make_golden_c1.py          (build container only)
The [2, 4096, D] queue is not stored: it is ffc.py:29-30's normalize(rand(2, Q, D)) drawn from a seeded generator that the test
re-creates; the final queue is pinned by a SHA-256 of its fp32 bytes (enqueue is a pure copy).  Writes tests/golden/c1_<case>.npz."""
import hashlib
import os
import sys

import numpy as np
import torch
import torch.nn.functional as F

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
from oracle import ref_shim  # noqa: E402
from make_golden import centers, make_batch  # noqa: E402

OUT = os.path.dirname(os.path.abspath(__file__))


def seeded_queue(Q, D, seed):
    q = torch.rand(2, Q, D, generator=torch.Generator().manual_seed(seed))
    return F.normalize(q, dim=2)                          # ffc.py:29-30


def c1_case(name, D, loss_type, margin, seed, Q=4096, B=64, n_ids=10000, steps=5, scale=32.0, noise=0.6, warm=1500):
    m = ref_shim.make_ffc(D, Q, scale, loss_type, margin, queue=seeded_queue(Q, D, seed))
    # a cold queue would make every early label an outlier: start from a partly filled LRU (identities 0..warm-1 resident, restore()
    # is part of the reference surface, lru.py:113-128) so that hits, misses and outliers all occur in the recorded steps
    m.lru.restore([(i, i) for i in range(warm)])
    gen = torch.Generator().manual_seed(seed + 1)
    cen = centers(n_ids, D, seed + 2)
    rec = dict(D=D, Q=Q, B=B, n_ids=n_ids, steps=steps, loss_type=loss_type, margin=margin, scale=scale, queue_seed=seed, warm=warm)
    for s in range(steps):
        x, y, xl, yl = make_batch(gen, cen, B, n_ids if s % 2 else warm + 200, noise)       # alternate mostly-known / mostly-new ids
        loss, dx, dy, tr = ref_shim.forward_backward(m, x, y, xl, yl)
        rec[f'x{s}'], rec[f'y{s}'] = x.numpy(), y.numpy()
        rec[f'xl{s}'], rec[f'yl{s}'] = np.array(xl), np.array(yl)
        rec[f'loss{s}'] = np.float64(loss)
        rec[f'dx{s}'], rec[f'dy{s}'] = dx.numpy(), dy.numpy()
        for pi, pn in enumerate(('rb', 'cm')):
            for k in ('rows', 'cols', 'labels', 'ones'):
                rec[f'{pn}_{k}{s}'] = np.array(tr[pi][k], dtype=np.int32)
        rec[f'lru{s}'] = np.array(m.lru.state_dict(), dtype=np.int32).reshape(-1, 2)
        rec[f'qpos{s}'] = np.array([m.queue_position_dict[i] for i in range(Q)], dtype=np.uint8)
    rec['queue_final_sha256'] = hashlib.sha256(m.queue.detach().contiguous().numpy().tobytes()).hexdigest()
    np.savez_compressed(os.path.join(OUT, f'c1_{name}.npz'), **rec)
    print(name, 'losses', [round(float(rec[f'loss{s}']), 4) for s in range(steps)],
          'outliers/step', [int((np.array(rec[f'cm_labels{s}']) < 0).sum()) for s in range(steps)])


if __name__ == '__main__':
    torch.set_num_threads(1)                              # serial index_put: "last duplicate wins" 
    c1_case('arc_d128', 128, 'Arc', 0.5, seed=31)
    c1_case('am_d512', 512, 'AM', 0.4, seed=32)
