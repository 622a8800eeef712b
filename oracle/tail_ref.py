"""CPU restatement of the reference backbones' tail (the hand-off into the FFC head), numpy float64.

TEST INFRASTRUCTURE ONLY: only ``tests/`` and ``__graft_entry__.smoke()`` may import this module; the product path
(``ffc_b200.tail``) never does.

What it restates:
  * resnet_arcface.py:99-101   ``self.features = nn.BatchNorm1d(feat_dim, eps=1e-05)``, weight fixed at 1, no grad
  * resnet_arcface.py:151      ``x = F.normalize(self.features(x))``
  * resnet_std.py:201-202      ``x = self.features(x); x = F.normalize(x)``
  * mobilefacenet_def.py:113-114   ``x = torch.flatten(x, 1); return F.normalize(x)``       (no BatchNorm1d)
The arithmetic behind those calls lives in PyTorch (torch 2.11 here), not under /root/reference: BatchNorm1d normalises a
2-D input per column with the batch mean and BIASED variance in train() and the running statistics in eval(), updates
``running = (1 - momentum) * running + momentum * batch`` with the UNBIASED variance, and F.normalize divides each row by
``max(||row||_2, 1e-12)``.  The backward below is the hand-derived adjoint of exactly that.

Parity is pinned against the reference's own backbones run in the build container: tests/golden/make_golden_tail.py hooks
the tail of ``iresnet50`` and ``MobileFaceNet`` during a real forward + backward and stores (x, p, dp, dx, dbias, running
statistics) in tests/golden/tail_*.npz; tests/test_oracle_tail.py checks this file against those fixtures.
"""
from __future__ import annotations

import numpy as np

NORM_EPS = 1e-12    # F.normalize default eps


def tail_forward(x, weight=None, bias=None, running_mean=None, running_var=None, training=True, eps=1e-5, momentum=0.1, bn=True):
    """Returns (p, cache, new_running_mean, new_running_var).  ``bn=False`` is the MobileFaceNet tail."""
    x = np.asarray(x, dtype=np.float64)
    B, D = x.shape
    new_rm, new_rv = running_mean, running_var
    if bn:
        w = np.ones(D) if weight is None else np.asarray(weight, dtype=np.float64)
        b = np.zeros(D) if bias is None else np.asarray(bias, dtype=np.float64)
        if training or running_mean is None:
            if B <= 1:
                raise ValueError('Expected more than 1 value per channel when training')
            mean = x.mean(axis=0)
            var = ((x - mean) ** 2).mean(axis=0)                       # biased
            if training and running_mean is not None:
                new_rm = (1 - momentum) * np.asarray(running_mean, dtype=np.float64) + momentum * mean
                new_rv = (1 - momentum) * np.asarray(running_var, dtype=np.float64) + momentum * var * B / (B - 1)
        else:
            mean = np.asarray(running_mean, dtype=np.float64)
            var = np.asarray(running_var, dtype=np.float64)
        invstd = 1.0 / np.sqrt(var + eps)
        xhat = (x - mean) * invstd
        y = xhat * w + b
    else:
        w = invstd = xhat = None
        y = x
    norm = np.sqrt((y * y).sum(axis=1, keepdims=True))
    denom = np.maximum(norm, NORM_EPS)
    p = y / denom
    cache = dict(bn=bn, batch_stats=bool(bn and (training or running_mean is None)), w=w, invstd=invstd, xhat=xhat, p=p, denom=denom,
                 clamped=norm < NORM_EPS)
    return p, cache, new_rm, new_rv


def tail_backward(dp, cache):
    """Returns (dx, dweight, dbias) -- dweight / dbias are None without BatchNorm1d."""
    dp = np.asarray(dp, dtype=np.float64)
    p, denom = cache['p'], cache['denom']
    # y / clamp_min(||y||, eps): the norm's gradient path exists only where the clamp is inactive
    s = np.where(cache['clamped'], 0.0, (p * dp).sum(axis=1, keepdims=True))
    dy = (dp - p * s) / denom
    if not cache['bn']:
        return dy, None, None
    xhat, w, invstd = cache['xhat'], cache['w'], cache['invstd']
    dbias = dy.sum(axis=0)
    dweight = (dy * xhat).sum(axis=0)
    if cache['batch_stats']:
        B = dy.shape[0]
        dx = w * invstd * (dy - dbias / B - xhat * dweight / B)
    else:
        dx = w * invstd * dy
    return dx, dweight, dbias
