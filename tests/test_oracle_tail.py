"""The tail oracle (oracle/tail_ref.py) against the fixtures captured from the reference's own backbones
(tests/golden/make_golden_tail.py: iresnet50 in train() / eval(), feat_dim 512 / 128, and MobileFaceNet)."""
import glob
import os

import numpy as np
import pytest

from oracle import tail_ref

CASES = sorted(os.path.basename(p)[5:-4] for p in glob.glob(os.path.join(os.path.dirname(__file__), 'golden', 'tail_*.npz')))
RTOL = 2e-5      # the fixtures are torch fp32 results; the oracle computes in float64


def close(a, b, what):
    a, b = np.asarray(a, dtype=np.float64), np.asarray(b, dtype=np.float64)
    err = np.abs(a - b).max()
    assert err <= RTOL * max(np.abs(b).max(), 1e-30), (what, err, np.abs(b).max())


def test_fixtures_present():
    assert {'ir50_train', 'ir50_eval', 'ir50_d128_train', 'mobile_d128'} <= set(CASES)


@pytest.mark.parametrize('case', CASES)
def test_oracle_matches_reference_backbone_tail(case, golden_dir):
    z = np.load(os.path.join(golden_dir, f'tail_{case}.npz'))
    bn = bool(z['bn'])
    for s in range(int(z['steps'])):
        if bn:
            p, cache, rm, rv = tail_ref.tail_forward(z[f'x{s}'], z['weight'], z['bias'], z[f'rm_in{s}'], z[f'rv_in{s}'], training=bool(z['training']),
                                                     eps=float(z['eps']), momentum=float(z['momentum']))
            close(rm, z[f'rm_out{s}'], 'running_mean')
            close(rv, z[f'rv_out{s}'], 'running_var')
        else:
            p, cache, _, _ = tail_ref.tail_forward(z[f'x{s}'], bn=False)
        close(p, z[f'p{s}'], 'p')
        np.testing.assert_allclose(np.linalg.norm(p, axis=1), 1.0, atol=1e-12)
        dx, dw, db = tail_ref.tail_backward(z[f'dp{s}'], cache)
        close(dx, z[f'dx{s}'], 'dx')
        if bn:
            close(db, z[f'dbias{s}'], 'dbias')
            assert not bool(z['weight_requires_grad'])          # resnet_arcface.py:101: the weight is frozen at 1


def test_oracle_backward_is_the_adjoint():
    """finite differences on the oracle itself (train-mode BatchNorm1d + normalise, and the clamped-row branch)"""
    rng = np.random.default_rng(0)
    B, D = 5, 7
    x = rng.normal(size=(B, D))
    w, b = rng.uniform(0.5, 1.5, D), rng.normal(size=D)
    dp = rng.normal(size=(B, D))
    for training in (True, False):
        rm, rv = rng.normal(size=D), rng.uniform(0.5, 2, D)

        def f(xx):
            return (tail_ref.tail_forward(xx, w, b, rm, rv, training=training)[0] * dp).sum()
        _, cache, _, _ = tail_ref.tail_forward(x, w, b, rm, rv, training=training)
        dx, _, _ = tail_ref.tail_backward(dp, cache)
        num = np.zeros_like(x)
        for i in range(B):
            for j in range(D):
                e = np.zeros_like(x)
                e[i, j] = 1e-6
                num[i, j] = (f(x + e) - f(x - e)) / 2e-6
        np.testing.assert_allclose(dx, num, rtol=1e-5, atol=1e-7)
    # an all-zero row is clamped: p = 0 / 1e-12 = 0 and dy = dp / 1e-12 (what torch autograd returns)
    x0 = np.zeros((2, 4))
    x0[1] = [3, 0, 4, 0]
    p, cache, _, _ = tail_ref.tail_forward(x0, bn=False)
    dx, _, _ = tail_ref.tail_backward(np.ones((2, 4)), cache)
    assert np.all(p[0] == 0) and np.allclose(dx[0], 1e12)
    np.testing.assert_allclose(p[1], [0.6, 0, 0.8, 0])
