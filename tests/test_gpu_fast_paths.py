"""GPU checks of the fused / merged entry points against the plain ones they replace (same inputs, same state):
ffc_head_pass_single == ffc_head_sweep + ffc_head_finalize, forward_pair == two head() calls, the sharded record path
(sweep_record -> gathered records -> finalize_gathered) emulated with two shard backends on ONE GPU == the unsharded
head, ffc_route_keys == a stable torch partition, ffc_queue_scatter_indexed == gather + ffc_queue_scatter."""
import ctypes as C

import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu


def _rel(a, b):
    return float((a - b).norm() / (b.norm() + 1e-30))


def _mk(D, Q, B, loss_type, margin, precision='bf16'):
    import ffc_b200
    return ffc_b200.FFCHead(D, Q, 32.0, loss_type, margin, precision=precision, max_batch=B, device=torch.device('cuda'))


def _batch(gen, B, D, n_ids):
    h = B // 2
    ids = torch.randperm(n_ids, generator=gen)[:h]
    xl = torch.cat([ids, torch.randint(0, n_ids, (B - h,), generator=gen)])
    yl = torch.cat([ids, torch.randint(0, n_ids, (B - h,), generator=gen)])
    return F.normalize(torch.randn(B, D, generator=gen)), F.normalize(torch.randn(B, D, generator=gen)), xl, yl


@pytest.mark.parametrize('D,Q,B,n_ids', [(512, 8192, 256, 9000), (128, 1000, 96, 3000)])
def test_pass_single_is_sweep_plus_finalize(D, Q, B, n_ids):
    """The fused reduce + coefficient + dEmb kernel against the three kernels it replaces (same sums, fp32 instead of fp64 scalars)."""
    from ffc_b200 import _capi
    from ffc_b200._capi import HeadPass, HeadStats, check
    dev = torch.device('cuda')
    torch.manual_seed(0)
    h = _mk(D, Q, B, 'Arc', 0.5)
    gen = torch.Generator().manual_seed(1)
    lib = _capi.lib()
    for step in range(3):
        x, y, xl, yl = _batch(gen, B, D, n_ids)
        loss, dp = h._pass(x.to(dev), y.to(dev), xl, yl, True)          # commit pass through ffc_head_pass_single
        # the same statistics through the two-call path on the state the pass left behind (label / ones / cmask of set 1 persist)
        st = h._sets[1]
        p32 = x.to(dev).contiguous()
        hp = HeadPass(p32.data_ptr(), h.queue.data_ptr(), h.queue_bf16.data_ptr(), st['label'].data_ptr(), st['ones_list'].data_ptr(),
                      st['n_ones'].data_ptr(), None, B)
        # cmask was cleared at the end of the pass: rebuild it from ones_list
        cm = torch.zeros_like(st['cmask'])
        n1 = int(st['n_ones'].item())
        for s in st['ones_list'][:n1].tolist():
            cm[s >> 5] |= (1 << (s & 31)) if (s & 31) < 31 else -(1 << 31)
        hp.cmask = cm.data_ptr()
        hs = HeadStats(*(h._stat_ptr(n, B) for n in ('lsum', 'osum', 'tgt', 'topv', 'topi')))
        loss2 = torch.empty((), device=dev)
        dp2 = torch.empty(B, D, device=dev)
        s = torch.cuda.current_stream().cuda_stream
        check(lib.ffc_head_sweep(h._h, C.byref(hp), C.byref(hs), s))
        check(lib.ffc_head_finalize(h._h, C.byref(hp), C.byref(hs), 1, loss2.data_ptr(), dp2.data_ptr(), s))
        # the fused kernel sums the partials in the same order; its per-row scalar part is fp32 (the three-kernel path: fp64)
        assert abs(float(loss) - float(loss2)) <= 2e-6 * abs(float(loss2)), (step, float(loss), float(loss2))
        assert _rel(dp, dp2) <= 2e-6, (step, _rel(dp, dp2))


def test_forward_pair_equals_two_passes():
    dev = torch.device('cuda')
    D, Q, B, n_ids = 256, 4096, 200, 5000
    torch.manual_seed(0)
    a, b = _mk(D, Q, B, 'AM', 0.4), _mk(D, Q, B, 'AM', 0.4)
    b.queue.copy_(a.queue)
    b._ensure()
    b.sync_mirror()
    gen = torch.Generator().manual_seed(2)
    for step in range(4):
        x, y, xl, yl = _batch(gen, B, D, n_ids)
        xd, yd = x.to(dev), y.to(dev)
        l2, dx = a._pass(xd, yd, xl, yl, False)
        l1, dy = a._pass(yd, xd, yl, xl, True)
        # pair entry: host labels on even steps (bookkeeping runs ahead on its own stream), device labels on odd steps
        lab = (xl, yl) if step % 2 == 0 else (xl.to(dev), yl.to(dev))
        lp, dxp, dyp = b.forward_pair(xd, yd, yd, xd, *lab)
        # bit-equal: every kernel of a pass sums in a fixed order (the LRU kernel sorts its `ones` segment by slot)
        assert float(lp) == float(l1 + l2)
        assert torch.equal(dxp, dx) and torch.equal(dyp, dy)
        assert a.lru.state_dict() == b.lru.state_dict()
        assert torch.equal(a.queue, b.queue) and torch.equal(a.qpos, b.qpos)


@pytest.mark.parametrize('R', [2, 4])
def test_record_path_two_shards_on_one_gpu(R):
    """R shard backends on one device, records concatenated by hand instead of an all-gather: loss and summed dEmb must match
    the unsharded head fed the same labels (global slots) and queue."""
    from ffc_b200.dist import CudaShardBackend
    from ffc_b200.ffc import hard_neg_k
    dev = torch.device('cuda')
    D, Q, n = 512, 8192, 384
    Ql = Q // R
    torch.manual_seed(0)
    full = _mk(D, Q, n, 'Arc', 0.5)
    full._ensure()
    gen = torch.Generator().manual_seed(3)
    shards = [CudaShardBackend(D, Ql, Q, r * Ql, n, 32.0, 'Arc', 0.5, hard_neg_k(Q), 'bf16', dev) for r in range(R)]
    for r, be in enumerate(shards):
        be.set_queue(full.queue[:, r * Ql:(r + 1) * Ql])
    p = F.normalize(torch.randn(n, D, generator=gen)).to(dev)
    # labels: 2/3 known (global slots), 1/3 outliers; a few `ones` slots per shard (their queue[1] rows differ from queue[0])
    label = torch.randint(0, Q, (n,), generator=gen).to(torch.int32)
    label[torch.rand(n, generator=gen) < 0.33] = -1
    ones_global = torch.randperm(Q, generator=gen)[:40].sort().values
    label[:10] = ones_global[:10].to(torch.int32)            # some targets inside the ones set
    label = label.to(dev)

    def set_ones(cmask, ones_list, n_ones, local):
        cmask.zero_()
        for j, s in enumerate(local):
            cmask[s >> 5] |= (1 << (s & 31)) if (s & 31) < 31 else -(1 << 31)
            ones_list[j] = s
        n_ones.fill_(len(local))

    # unsharded reference through the two-call API
    from ffc_b200 import _capi
    from ffc_b200._capi import HeadPass, HeadStats, check
    lib = _capi.lib()
    st = full._sets[0]
    set_ones(st['cmask'], st['ones_list'], st['n_ones'], ones_global.tolist())
    hp = HeadPass(p.data_ptr(), full.queue.data_ptr(), full.queue_bf16.data_ptr(), label.data_ptr(), st['ones_list'].data_ptr(),
                  st['n_ones'].data_ptr(), st['cmask'].data_ptr(), n)
    hs = HeadStats(*(full._stat_ptr(nm, n) for nm in ('lsum', 'osum', 'tgt', 'topv', 'topi')))
    loss_ref = torch.empty((), device=dev)
    dp_ref = torch.empty(n, D, device=dev)
    s = torch.cuda.current_stream().cuda_stream
    check(lib.ffc_head_pass_single(full._h, C.byref(hp), C.byref(hs), loss_ref.data_ptr(), dp_ref.data_ptr(), s))

    recs = [be.new_records(n, R) for be in shards]
    for r, be in enumerate(shards):
        be.use_set(0)
        local = [int(g) - r * Ql for g in ones_global.tolist() if r * Ql <= int(g) < (r + 1) * Ql]
        set_ones(be.cmask, be.ones_list, be.n_ones, local)
        be.sweep_record(p, label, recs[r])
    gathered = torch.stack([rc['own'] for rc in recs])
    dp_sum = torch.zeros(n, D, device=dev)
    losses = []
    for r, be in enumerate(shards):
        recs[r]['all'].copy_(gathered)
        loss, dp = be.finalize_gathered(p, label, recs[r], R)
        losses.append(float(loss))
        dp_sum += dp
    assert all(abs(l - losses[0]) <= 1e-6 * abs(losses[0]) for l in losses)
    assert abs(losses[0] - float(loss_ref)) <= 2e-5 * abs(float(loss_ref)), (losses[0], float(loss_ref))
    assert _rel(dp_sum, dp_ref) <= 2e-5, _rel(dp_sum, dp_ref)


def test_route_keys_and_indexed_scatter():
    from ffc_b200 import _capi
    from ffc_b200._capi import check
    lib = _capi.lib()
    dev = torch.device('cuda')
    gen = torch.Generator().manual_seed(4)
    s = torch.cuda.current_stream().cuda_stream
    for n, R, rank in ((8192, 8, 3), (2048, 2, 1), (1000, 4, 0), (5, 3, 2), (3000, 1, 0)):
        keys = torch.randint(-50, 1 << 40, (n,), generator=gen).to(dev)
        keys_c = torch.empty_like(keys)
        order = torch.empty(n, dtype=torch.int32, device=dev)
        n_mine = torch.empty(1, dtype=torch.int32, device=dev)
        check(lib.ffc_route_keys(keys.data_ptr(), n, R, rank, keys_c.data_ptr(), order.data_ptr(), n_mine.data_ptr(), s))
        mine = torch.remainder(keys, R) == rank
        want = torch.argsort((~mine).to(torch.int8), stable=True)
        assert int(n_mine) == int(mine.sum())
        assert torch.equal(order.long(), want) and torch.equal(keys_c, keys[want])
    # indexed scatter == gather + scatter
    Q, D, B = 64, 16, 12
    q1 = torch.randn(2, Q, D, device=dev)
    q2 = q1.clone()
    h1 = torch.zeros(2, Q, D, dtype=torch.bfloat16, device=dev)
    h2 = h1.clone()
    rows = torch.randint(0, 2, (B,), generator=gen).to(torch.int32).to(dev)
    cols = torch.randint(0, 8, (B,), generator=gen).to(torch.int32).to(dev)      # duplicates on purpose
    cols[-2:] = -1
    g_all = torch.randn(40, D, device=dev)
    src = torch.randperm(40, generator=gen)[:B].to(torch.int32).to(dev)
    u1, u2 = torch.zeros(B, D, device=dev), torch.zeros(B, D, device=dev)
    check(lib.ffc_queue_scatter_indexed(q1.data_ptr(), h1.data_ptr(), rows.data_ptr(), cols.data_ptr(), g_all.data_ptr(), src.data_ptr(), B, Q, D,
                                        u1.data_ptr(), s))
    gc = g_all[src.long()].contiguous()
    check(lib.ffc_queue_scatter(q2.data_ptr(), h2.data_ptr(), rows.data_ptr(), cols.data_ptr(), gc.data_ptr(), B, Q, D, u2.data_ptr(), s))
    assert torch.equal(q1, q2) and torch.equal(h1, h2) and torch.equal(u1, u2)


@pytest.mark.parametrize('D,Q,B,n_ids', [(256, 4096, 200, 5000), (512, 65536, 512, 100000)])
def test_pass_is_run_to_run_deterministic(D, Q, B, n_ids):
    """A rollback pass leaves the head untouched, so repeating it must reproduce loss and dEmb bit for bit (hits, fresh inserts,
    `ones`, outliers and the hard-negative top-k all exercised)."""
    dev = torch.device('cuda')
    torch.manual_seed(0)
    h = _mk(D, Q, B, 'Arc', 0.5)
    h._ensure()
    n0 = min(Q, n_ids) * 3 // 4
    h.lru.restore_arrays(torch.arange(n0, dtype=torch.int64), torch.arange(n0, dtype=torch.int32))
    gen = torch.Generator().manual_seed(1)
    x, y, xl, yl = _batch(gen, B, D, n_ids)
    x, y = x.to(dev), y.to(dev)
    l0, d0 = h._pass(x, y, xl, yl, False)
    l0, d0 = float(l0), d0.clone()
    for _ in range(5):
        l, d = h._pass(x, y, xl, yl, False)
        assert float(l) == l0 and torch.equal(d, d0)
