"""Pin oracle/head_ref.py against fixtures generated from the reference FFC head (tests/golden/ffc_*.npz)."""
import glob
import os

import numpy as np
import pytest
import torch

from oracle.head_ref import HeadOracle, hard_neg_k

CASES = sorted(os.path.basename(p)[4:-4] for p in glob.glob(os.path.join(os.path.dirname(__file__), 'golden', 'ffc_*.npz')))


def test_hard_neg_k():
    # ffc.py:48, probed values from SURVEY.md section 3.4
    assert hard_neg_k(1000) == 3 and hard_neg_k(16384) == 3 and hard_neg_k(50000) == 10 and hard_neg_k(1 << 20) == 10
    assert hard_neg_k(25000) == 5


@pytest.mark.parametrize('name', CASES)
@pytest.mark.parametrize('dtype', [torch.float32, torch.float64])
def test_head_matches_reference(golden_dir, name, dtype):
    z = np.load(os.path.join(golden_dir, f'ffc_{name}.npz'))
    D, Q = int(z['D']), int(z['Q'])
    o = HeadOracle(D, Q, float(z['scale']), str(z['loss_type']), float(z['margin']),
                   queue=torch.from_numpy(z['queue0']), dtype=dtype)
    for s in range(int(z['steps'])):
        x = torch.from_numpy(z[f'x{s}']).to(dtype).requires_grad_(True)
        y = torch.from_numpy(z[f'y{s}']).to(dtype).requires_grad_(True)
        loss = o.forward(x, y, z[f'xl{s}'].tolist(), z[f'yl{s}'].tolist())
        loss.backward()
        rb, cm = o.trace[-2], o.trace[-1]
        for tr, pn in ((rb, 'rb'), (cm, 'cm')):
            for k in ('rows', 'cols', 'labels', 'ones'):
                assert tr[k] == z[f'{pn}_{k}{s}'].tolist(), (s, pn, k)
        assert [list(kv) for kv in o.lru.state_dict()] == z[f'lru{s}'].tolist()
        assert o.qpos == z[f'qpos{s}'].tolist()
        ref = float(z[f'loss{s}'])
        assert abs(float(loss) - ref) <= 2e-5 * abs(ref), (s, float(loss), ref)
        for got, want in ((x.grad, z[f'dx{s}']), (y.grad, z[f'dy{s}'])):
            want = torch.from_numpy(want).to(dtype)
            assert (got - want).norm() <= 2e-5 * want.norm() + 1e-7
    assert np.abs(o.queue.float().numpy() - z['queue_final']).max() < 1e-6
