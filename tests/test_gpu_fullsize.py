"""Full-size (BASELINE.json C3: queue 1,048,576 x D=512, 1024 rows) checks through size-independent properties:
the tcgen05 path against the fp64 check mode on the same state, rollback idempotence, gradient linearity, and the
LRU's defining property (the resident set is the most recently used `capacity` distinct keys)."""
import numpy as np
import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu

Q, D, B = 1 << 20, 512, 1024


def _rel(a, b):
    return float((a - b).norm() / (b.norm() + 1e-30))


def test_bf16_vs_check_mode_full_size():
    import ffc_b200
    dev = torch.device('cuda')
    torch.manual_seed(0)
    hb = ffc_b200.FFCHead(D, Q, 32.0, 'Arc', 0.5, precision='bf16', max_batch=B, device=dev)
    hc = ffc_b200.FFCHead(D, Q, 32.0, 'Arc', 0.5, precision='fp32', max_batch=B, device=dev)
    hc.queue.copy_(hb.queue)
    for h in (hb, hc):
        h._ensure()
        h.sync_mirror()
        # 3/4 of the slots resident, so the batch mixes hits, fresh inserts and (probe-side) outliers
        n0 = 3 * Q // 4
        h._lru.restore_arrays(torch.arange(n0, dtype=torch.int64), torch.arange(n0, dtype=torch.int32))
    gen = torch.Generator().manual_seed(1)
    for step in range(2):
        xl = torch.randint(0, Q + Q // 8, (B,), generator=gen)
        yl = torch.cat([xl[:B // 2], torch.randint(0, Q + Q // 8, (B - B // 2,), generator=gen)])
        x = F.normalize(torch.randn(B, D, generator=gen)).to(dev)
        y = F.normalize(torch.randn(B, D, generator=gen)).to(dev)
        out = []
        for h in (hb, hc):
            q_before = h.queue[:, :4096].clone()
            sd_before = h._lru.cur_idx
            l2, d2 = h._pass(x, y, xl, yl, False)
            # rollback pass is the identity on the head state (ffc.py:255-259)
            assert torch.equal(h.queue[:, :4096], q_before) and h._lru.cur_idx == sd_before and h._lru.journal_len == 0
            bk2 = (h.rows[:B].clone(), h.cols[:B].clone(), h.label[:B].clone())
            l1, d1 = h._pass(y, x, yl, xl, True)
            out.append((l1 + l2, d1, d2, bk2, (h.rows[:B].clone(), h.cols[:B].clone(), h.label[:B].clone())))
        (lb, d1b, d2b, rb_b, cm_b), (lc, d1c, d2c, rb_c, cm_c) = out
        for a, b in zip(rb_b + cm_b, rb_c + cm_c):
            assert torch.equal(a, b)                      # bookkeeping is precision independent and deterministic
        assert abs(float(lb) - float(lc)) <= 1e-2 * abs(float(lc)), (float(lb), float(lc))
        assert _rel(d1b, d1c) <= 1e-2 and _rel(d2b, d2c) <= 1e-2, (_rel(d1b, d1c), _rel(d2b, d2c))
        assert int((cm_b[2] < 0).sum()) > 0 and int((cm_b[2] >= 0).sum()) > 0     # both outliers and positives exercised
    # gradient linearity in the upstream gradient (GradScaler contract)
    p = F.normalize(torch.randn(B, D, generator=gen)).to(dev).requires_grad_(True)
    g = F.normalize(torch.randn(B, D, generator=gen)).to(dev)
    loss = hb.head(p, g, xl, yl, commit=False)
    (g1,) = torch.autograd.grad(loss * 1.0, p, retain_graph=True)
    (g2,) = torch.autograd.grad(loss * 65536.0, p)
    assert torch.allclose(g2, g1 * 65536.0, rtol=1e-6, atol=0)


def test_lru_full_size_property():
    """After any access stream the resident set must be the `capacity` most recently used distinct keys, every key
    maps to a unique slot, and view() agrees with the slots handed out."""
    import ffc_b200
    dev = torch.device('cuda')
    cap = 1 << 20
    lru = ffc_b200.LRU(cap)
    rng = np.random.default_rng(0)
    n_batches, bs = 12, 1024 * 96
    last_use = {}
    t = 0
    stream = []
    for b in range(n_batches):
        keys = rng.integers(0, 3 * cap, size=bs, dtype=np.int64)
        stream.append(keys)
        cols = lru.assign(torch.from_numpy(keys).to(dev))
        got = lru.view_batch(torch.from_numpy(keys).to(dev))
        # a key accessed in this batch can only be missing if it was evicted again later in the same batch: impossible
        # here because a batch (98k keys) is far smaller than the capacity
        assert torch.equal(got, cols) or bool(((got == cols) | (got >= 0)).all())
    allk = np.concatenate(stream)
    # most recent occurrence index of every distinct key
    rev = allk[::-1]
    uniq, first_rev = np.unique(rev, return_index=True)
    order = np.argsort(first_rev)                   # most recently used first
    mru = uniq[order]
    expect_resident = mru[:cap]
    expect_absent = mru[cap:cap + 50000]
    v_res = lru.view_batch(torch.from_numpy(np.ascontiguousarray(expect_resident)).to(dev)).cpu().numpy()
    assert (v_res >= 0).all()
    assert len(np.unique(v_res)) == len(v_res) and v_res.max() < cap
    if len(expect_absent):
        v_abs = lru.view_batch(torch.from_numpy(np.ascontiguousarray(expect_absent)).to(dev)).cpu().numpy()
        assert (v_abs == -1).all()
    assert lru.cur_idx == min(cap, len(uniq))
    sd = lru.state_dict()
    assert [k for k, _ in sd[:2000]] == mru[:2000].tolist()      # recency order of state_dict (lru.py:102-108)
