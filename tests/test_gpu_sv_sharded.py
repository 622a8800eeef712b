"""SV loss (ffc.py:116-138) on a column-sharded queue, emulated with R shard backends on ONE GPU: ffc_head_prep on every shard ->
target cosines summed over the shards (the all-reduce of ffc_b200/dist.py) -> ffc_head_sweep_prepared -> denominators summed, top-k
candidates gathered -> ffc_head_finalize; loss and summed dEmb must match the unsharded head (ffc_head_sweep + ffc_head_finalize)
fed the same global labels and queue.  Also: prep + sweep_prepared == sweep for AM / Arc on one shard."""
import ctypes as C

import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu


def _rel(a, b):
    return float((a - b).norm() / (b.norm() + 1e-30))


def _set_ones(cmask, ones_list, n_ones, local):
    cmask.zero_()
    for j, s in enumerate(local):
        cmask[s >> 5] |= (1 << (s & 31)) if (s & 31) < 31 else -(1 << 31)
        ones_list[j] = s
    n_ones.fill_(len(local))


def _inputs(D, Q, n, seed):
    gen = torch.Generator().manual_seed(seed)
    dev = torch.device('cuda')
    # clustered probes: each known row sits near its target prototype, so SV's hard-example branch (cos > gt - margin) is
    # taken for a handful of columns per row and the final_gt branch (gt > margin) for most rows
    label = torch.randint(0, Q, (n,), generator=gen).to(torch.int32)
    label[torch.rand(n, generator=gen) < 0.3] = -1
    ones_global = torch.randperm(Q, generator=gen)[:40].sort().values
    label[:10] = ones_global[:10].to(torch.int32)
    return gen, dev, label, ones_global


@pytest.mark.parametrize('precision,tol', [('fp32', 1e-5), ('bf16', 1e-4)])
@pytest.mark.parametrize('R', [2, 4])
def test_sv_shards_on_one_gpu(R, precision, tol):
    import ffc_b200
    from ffc_b200 import _capi
    from ffc_b200._capi import HeadPass, HeadStats, check
    from ffc_b200.dist import CudaShardBackend
    from ffc_b200.ffc import hard_neg_k
    D, Q, n = 128, 4096, 200
    Ql = Q // R
    torch.manual_seed(0)
    gen, dev, label, ones_global = _inputs(D, Q, n, 13)
    full = ffc_b200.FFCHead(D, Q, 32.0, 'SV', 0.4, precision=precision, max_batch=n, device=dev)
    # zero-mean prototypes (the default all-positive initialisation, ffc.py:29-30, has every pair of rows at cos ~ 0.75: every
    # column would be a hard example); with these, cos(p, other) ~ N(0, 1/sqrt(D)) and gt - margin cuts ~2 sigma out
    with torch.no_grad():
        full.queue.copy_(F.normalize(torch.randn(2, Q, D, generator=gen), dim=2))
    full._ensure()
    full.sync_mirror()
    known = label >= 0
    p = torch.randn(n, D, generator=gen)
    proto = full.queue[0].cpu()
    p[known] = proto[label[known].long()] + 0.12 * p[known]
    p = F.normalize(p).to(dev)
    label = label.to(dev)
    lib = _capi.lib()
    s = torch.cuda.current_stream().cuda_stream

    st = full._sets[0]
    _set_ones(st['cmask'], st['ones_list'], st['n_ones'], ones_global.tolist())
    hp = HeadPass(p.data_ptr(), full.queue.data_ptr(), full.queue_bf16.data_ptr(), label.data_ptr(), st['ones_list'].data_ptr(),
                  st['n_ones'].data_ptr(), st['cmask'].data_ptr(), n)
    hs = HeadStats(*(full._stat_ptr(nm, n) for nm in ('lsum', 'osum', 'tgt', 'topv', 'topi')))
    loss_ref = torch.empty((), device=dev)
    dp_ref = torch.empty(n, D, device=dev)
    check(lib.ffc_head_sweep(full._h, C.byref(hp), C.byref(hs), s))
    check(lib.ffc_head_finalize(full._h, C.byref(hp), C.byref(hs), 1, loss_ref.data_ptr(), dp_ref.data_ptr(), s))
    # the test is only meaningful if SV's hard-example branch fires: some non-target cosine above gt - margin
    cos = p @ full.queue[0].t()
    gt = cos[known.to(dev)].gather(1, label[known.to(dev)].long().view(-1, 1))
    n_hard = int(((cos[known.to(dev)] > gt - 0.4).sum(1) - 1).clamp_min(0).sum())
    assert 0 < n_hard < 0.2 * int(known.sum()) * Q, n_hard            # both branches of ffc.py:121-125 are exercised

    shards = [CudaShardBackend(D, Ql, Q, r * Ql, n, 32.0, 'SV', 0.4, hard_neg_k(Q), precision, dev) for r in range(R)]
    stats = []
    for r, be in enumerate(shards):
        be.set_queue(full.queue[:, r * Ql:(r + 1) * Ql])
        be.use_set(0)
        _set_ones(be.cmask, be.ones_list, be.n_ones, [int(g) - r * Ql for g in ones_global.tolist() if r * Ql <= int(g) < (r + 1) * Ql])
        stats.append(be.new_stats(n, R))
        be.prep(p, label, stats[r], r)
    tgt_sum = sum(stt['red'][4:] for stt in stats)                  # dist.all_reduce(st['red'][4:])
    for r, be in enumerate(shards):
        stats[r]['red'][4:].copy_(tgt_sum)
        be.sweep_prepared(p, label, stats[r], r)
    lsum = sum(stt['red'][:4] for stt in stats)                     # dist.all_reduce(st['red'][:4])
    topv = torch.stack([stats[r]['topv'][r] for r in range(R)])    # all-gather of each rank's own candidate set
    topi = torch.stack([stats[r]['topi'][r] for r in range(R)])
    dp_sum = torch.zeros(n, D, device=dev)
    losses = []
    for r, be in enumerate(shards):
        stats[r]['red'][:4].copy_(lsum)
        stats[r]['topv'].copy_(topv)
        stats[r]['topi'].copy_(topi)
        loss, dp = be.finalize(p, label, stats[r], R)
        losses.append(float(loss))
        dp_sum += dp                                                # reduce-scatter SUM
    assert all(abs(l - losses[0]) <= 1e-6 * abs(losses[0]) for l in losses), losses
    assert abs(losses[0] - float(loss_ref)) <= tol * abs(float(loss_ref)), (losses[0], float(loss_ref), n_hard)
    assert _rel(dp_sum, dp_ref) <= 2 * tol, (_rel(dp_sum, dp_ref), n_hard)

    # wrong pairing is an error, not a silent sweep with stale thresholds
    with pytest.raises(ffc_b200.FFCError):
        shards[0].sweep_prepared(p, label, stats[0], 0)


@pytest.mark.parametrize('loss_type,margin', [('AM', 0.4), ('Arc', 0.5), ('SV', 0.4)])
def test_prep_plus_sweep_prepared_equals_sweep(loss_type, margin):
    """one shard covering the whole queue: the two-step form gives bit-identical statistics"""
    from ffc_b200.dist import CudaShardBackend
    from ffc_b200.ffc import hard_neg_k
    D, Q, n = 128, 2048, 160
    gen, dev, label, ones_global = _inputs(D, Q, n, 17)
    torch.manual_seed(1)
    be = CudaShardBackend(D, Q, Q, 0, n, 32.0, loss_type, margin, hard_neg_k(Q), 'bf16', dev)
    p = F.normalize(torch.randn(n, D, generator=gen)).to(dev)
    label = label.to(dev)
    be.use_set(0)
    _set_ones(be.cmask, be.ones_list, be.n_ones, ones_global.tolist())
    a, b = be.new_stats(n, 1), be.new_stats(n, 1)
    be.sweep(p, label, a, 0)
    be.prep(p, label, b, 0)
    be.sweep_prepared(p, label, b, 0)
    assert torch.equal(a['red'], b['red'])
    la, da = be.finalize(p, label, a, 1)
    lb, db = be.finalize(p, label, b, 1)
    assert torch.equal(la, lb) and torch.equal(da, db)
