#!/usr/bin/env python
"""Small single-GPU run of the sharded record path (R shard backends on one device) against the unsharded head, for
compute-sanitizer (memcheck / initcheck / racecheck):   compute-sanitizer --tool initcheck python tools/sanitize_repro.py"""
import ctypes as C
import os
import sys

R_ = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [R_, os.path.join(R_, 'very-large-scale-face-recognition_b200')]
import torch
import torch.nn.functional as F
import ffc_b200
from ffc_b200 import _capi
from ffc_b200._capi import HeadPass, HeadStats, check
from ffc_b200.dist import CudaShardBackend
from ffc_b200.ffc import hard_neg_k

dev = torch.device('cuda')
D, Q, n, R = int(os.environ.get('REPRO_D', 128)), int(os.environ.get('REPRO_Q', 1024)), int(os.environ.get('REPRO_N', 192)), 2
pad = os.environ.get('REPRO_PAD')
if pad:
    junk = torch.full((int(pad),), float('nan'), device=dev)      # shift the allocation layout / poison
    del junk
Ql = Q // R
torch.manual_seed(0)
full = ffc_b200.FFCHead(D, Q, 32.0, 'Arc', 0.5, precision='bf16', max_batch=n, device=dev)
full._ensure()
gen = torch.Generator().manual_seed(3)
shards = [CudaShardBackend(D, Ql, Q, r * Ql, n, 32.0, 'Arc', 0.5, hard_neg_k(Q), 'bf16', dev) for r in range(R)]
for r, be in enumerate(shards):
    be.set_queue(full.queue[:, r * Ql:(r + 1) * Ql])
lib = _capi.lib()
for trial in range(3):
    p = F.normalize(torch.randn(n, D, generator=gen)).to(dev)
    label = torch.randint(0, Q, (n,), generator=gen).to(torch.int32)
    label[torch.rand(n, generator=gen) < 0.4] = -1
    ones_global = torch.randperm(Q, generator=gen)[:int(os.environ.get('REPRO_ONES', 24))].sort().values
    label[:6] = ones_global[:6].to(torch.int32)
    label = label.to(dev)

    def set_ones(cmask, ones_list, n_ones, local):
        cmask.zero_()
        for j, s in enumerate(local):
            cmask[s >> 5] |= (1 << (s & 31)) if (s & 31) < 31 else -(1 << 31)
            ones_list[j] = s
        n_ones.fill_(len(local))
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
    row_err = (dp_sum - dp_ref).norm(dim=1) / (dp_ref.norm(dim=1) + 1e-30)
    bad = torch.nonzero(row_err > 1e-3).flatten().tolist()
    if bad:
        # per-rank record candidates of the first bad row against a torch recomputation
        b = bad[0]
        k = hard_neg_k(Q)
        pb = p[b].to(torch.bfloat16).float()
        for r, be in enumerate(shards):
            w = recs[r]['words']
            rec = recs[r]['own'][0]
            topv = rec[8 * n:8 * n + 3 * n * k].view(3, n, k)[:, b]
            topi = rec[8 * n + 3 * n * k:].view(torch.int32).view(3, n, k)[:, b]
            n1 = int(be.n_ones)
            ol = be.ones_list[:n1].long()
            c0 = be.queue_bf16[0].float() @ pb
            c1 = be.queue_bf16[1].float() @ pb
            cm = c0.clone(); cm[ol] = -9
            print(f'  row {b} rank {r}: record common {topv[0].tolist()} {topi[0].tolist()} | torch {torch.topk(cm, k).values.tolist()} {(torch.topk(cm, k).indices + r * Ql).tolist()}')
            for l, cc in ((1, c0), (2, c1)):
                tv = torch.topk(cc[ol], min(k, n1))
                print(f'           side{l - 1} record {topv[l].tolist()} {topi[l].tolist()} | torch {tv.values.tolist()} {(ol[tv.indices] + r * Ql).tolist()} (n_ones {n1})')
    # every outlier row's recorded candidates against a torch top-k over the same bf16 operands
    k = hard_neg_k(Q)
    p16 = p.to(torch.bfloat16).float()
    n_mis = 0
    for r, be in enumerate(shards):
        rec = recs[r]['own'][0]
        topv = rec[8 * n:8 * n + 3 * n * k].view(3, n, k)
        n1 = int(be.n_ones)
        ol = be.ones_list[:n1].long()
        c0 = p16 @ be.queue_bf16[0].float().t()
        c1 = p16 @ be.queue_bf16[1].float().t()
        cm = c0.clone()
        cm[:, ol] = -9.0
        want = [torch.topk(cm, k, dim=1).values, torch.topk(c0[:, ol], min(k, n1), dim=1).values, torch.topk(c1[:, ol], min(k, n1), dim=1).values]
        outl = (label < 0)
        for sidx in range(3):
            got = topv[sidx][outl][:, :want[sidx].shape[1]]
            w_ = want[sidx][outl].clamp_min(0)          # only positive cosines are kept (the rest cannot contribute)
            g_ = torch.where(torch.isinf(got), torch.zeros_like(got), got).clamp_min(0)
            d = (g_ - w_).abs().max(dim=1).values
            badr = torch.nonzero(d > 2e-3).flatten()
            n_mis += int(badr.numel())
            if badr.numel():
                i = int(badr[0])
                print(f'  rank {r} set {sidx}: {int(badr.numel())} outlier rows differ; first: got {g_[i].tolist()} want {w_[i].tolist()} (n_ones {n1})')
    print(f'  top-k candidate mismatches vs torch: {n_mis}')
    print(f'trial {trial}: loss {losses} vs {float(loss_ref):.6f}; max row err {float(row_err.max()):.3e}; bad rows {bad[:10]} labels {[int(label[b]) for b in bad[:10]]}', flush=True)
