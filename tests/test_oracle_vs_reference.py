"""Live cross-checks of the oracle against the UNMODIFIED reference (/root/reference), randomised.  Only where the reference tree exists
(the build container); skipped on the GPU box, where the committed fixtures (tests/golden/) stand in for it.

  * oracle/lru_ref.py  vs  reference lru.py:   random op streams (get / try_get / rollback / view / contains / state_dict / restore)
  * oracle/head_ref.py vs  reference ffc.py:   random small FFC configurations, several FFC.forward + backward steps each -- integer
                                               bookkeeping bit-exact, loss / gradients / queue within fp32 tolerance
  * oracle/tail_ref.py vs  torch's BatchNorm1d + F.normalize (what resnet_arcface.py:151 calls), random shapes and modes
"""
import random

import numpy as np
import pytest
import torch
import torch.nn.functional as F

from oracle import ref_shim, tail_ref
from oracle.head_ref import HeadOracle
from oracle.lru_ref import LRU

pytestmark = pytest.mark.skipif(not ref_shim.available(), reason='reference tree not present (GPU box): the golden fixtures cover this')


@pytest.mark.parametrize('seed', range(6))
def test_lru_random_streams(seed):
    _, lru_mod = ref_shim.load()
    rng = random.Random(seed)
    cap = rng.choice([1, 2, 3, 5, 8, 13, 32])
    universe = rng.choice([cap, cap + 1, 2 * cap + 3, 10 * cap])
    ref, ora = lru_mod.LRU(cap), LRU(cap)
    pending = 0
    for step in range(1500):
        r, key = rng.random(), rng.randrange(universe)
        if r < 0.40 and pending == 0:
            assert ora.get(key) == ref.get(key)
        elif r < 0.70:
            assert ora.try_get(key) == ref.try_get(key)
            pending += 1
        elif r < 0.80 and pending:
            n = rng.randrange(1, pending + 1)
            ref.rollback_steps(n), ora.rollback_steps(n)
            pending -= n
        elif r < 0.85 and pending:
            ref.rollback_one_step(), ora.rollback_one_step()
            pending -= 1
        elif r < 0.92:
            assert ora.view(key) == ref.view(key) and (key in ora) == (key in ref)
        elif r < 0.97:
            assert ora.state_dict() == ref.state_dict() and ora.cur_idx == ref.cur_idx and sorted(ora.keys()) == sorted(ref.keys())
            assert list(ora) == list(ref)
        elif pending == 0:
            # restore round trip into fresh caches (lru.py:113-128)
            kvs = ref.state_dict()
            ref, ora = lru_mod.LRU(cap), LRU(cap)
            ref.restore(kvs), ora.restore(kvs)
    assert ora.state_dict() == ref.state_dict() and ora.cur_idx == ref.cur_idx


@pytest.mark.parametrize('seed', range(4))
@pytest.mark.parametrize('loss_type', ['AM', 'Arc', 'SV'])
def test_head_random_configs(seed, loss_type):
    rng = random.Random(100 * seed + len(loss_type))
    D = rng.choice([4, 16, 32])
    Q = rng.choice([6, 17, 40])
    B = rng.choice([4, 6, 10])
    n_ids = rng.choice([Q // 2 + 2, Q + 3, 3 * Q])          # all-hit, mixed and eviction-heavy regimes
    margin = 0.5 if loss_type == 'Arc' else 0.4
    torch.manual_seed(seed)
    m = ref_shim.make_ffc(D, Q, 32.0, loss_type, margin)
    o = HeadOracle(D, Q, 32.0, loss_type, margin, queue=m.queue.detach().clone(), dtype=torch.float32)
    gen = torch.Generator().manual_seed(seed + 50)
    for step in range(5):
        ids = torch.randperm(n_ids, generator=gen)[:B // 2]
        xl = torch.cat([ids, torch.randint(0, n_ids, (B - B // 2,), generator=gen)])
        yl = torch.cat([ids, torch.randint(0, n_ids, (B - B // 2,), generator=gen)])
        x = F.normalize(torch.randn(B, D, generator=gen))
        y = F.normalize(torch.randn(B, D, generator=gen))
        loss_ref, dx_ref, dy_ref, trace = ref_shim.forward_backward(m, x, y, xl.tolist(), yl.tolist())
        xo, yo = x.clone().requires_grad_(True), y.clone().requires_grad_(True)
        loss = o.forward(xo, yo, xl.tolist(), yl.tolist())
        if torch.is_tensor(loss) and loss.requires_grad:
            loss.backward()
        for got, want in zip(o.trace[-2:], trace):
            for k in ('rows', 'cols', 'labels', 'ones'):
                assert got[k] == want[k], (step, k)
        assert o.lru.state_dict() == m.lru.state_dict() and o.lru.cur_idx == m.lru.cur_idx
        assert o.qpos == [m.queue_position_dict[i] for i in range(Q)]
        if np.isfinite(loss_ref):       # Arc: |cos_t| can reach 1 (sqrt(1 - 1) has an infinite slope, ffc.py:101): NaN on both sides
            assert abs(float(loss) - loss_ref) <= 2e-5 * abs(loss_ref) + 1e-6, (step, float(loss), loss_ref)
            for got, want in ((xo.grad, dx_ref), (yo.grad, dy_ref)):
                got = torch.zeros_like(want) if got is None else got
                assert (got - want).norm() <= 5e-5 * want.norm() + 1e-6, step
        else:
            assert not np.isfinite(float(loss))
        assert torch.allclose(o.queue.float(), m.queue.float(), atol=1e-6)


@pytest.mark.parametrize('seed', range(5))
def test_tail_against_torch(seed):
    rng = np.random.default_rng(seed)
    B, D = int(rng.integers(2, 40)), int(rng.integers(1, 70))
    for training in (True, False):
        for affine in (True, False):
            bn = torch.nn.BatchNorm1d(D, eps=1e-05, affine=affine).double()
            with torch.no_grad():
                bn.running_mean.normal_(0, 0.3)
                bn.running_var.uniform_(0.5, 2.0)
                if affine:
                    bn.weight.uniform_(0.5, 1.5)
                    bn.bias.normal_(0, 0.3)
            bn.train(training)
            rm0, rv0 = bn.running_mean.numpy().copy(), bn.running_var.numpy().copy()
            x = torch.tensor(rng.normal(0.2, 1.5, size=(B, D)), requires_grad=True)
            dp = torch.tensor(rng.normal(size=(B, D)))
            p = F.normalize(bn(x))
            p.backward(dp)
            w = bn.weight.detach().numpy() if affine else None
            b = bn.bias.detach().numpy() if affine else None
            po, cache, rm, rv = tail_ref.tail_forward(x.detach().numpy(), w, b, rm0, rv0, training=training)
            dx, dw, db = tail_ref.tail_backward(dp.numpy(), cache)
            np.testing.assert_allclose(po, p.detach().numpy(), rtol=1e-10, atol=1e-12)
            np.testing.assert_allclose(dx, x.grad.numpy(), rtol=1e-8, atol=1e-11)
            np.testing.assert_allclose(rm, bn.running_mean.numpy(), rtol=1e-12, atol=1e-14)
            np.testing.assert_allclose(rv, bn.running_var.numpy(), rtol=1e-12, atol=1e-14)
            if affine:
                np.testing.assert_allclose(dw, bn.weight.grad.numpy(), rtol=1e-8, atol=1e-11)
                np.testing.assert_allclose(db, bn.bias.grad.numpy(), rtol=1e-8, atol=1e-11)
