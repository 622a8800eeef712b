"""TEST-ONLY torch/CPU stand-in for one rank's shard of the sharded FFC head (ffc_b200.dist.CudaShardBackend).

It restates, in fp64 torch, what the CUDA kernels compute per shard -- bookkeeping through the oracle LRU, the sweep
statistics (softmax denominators relative to the fixed max M = scale, sum_j p~_j W_j, target cosines, top-k
candidates) and the finalize formulas of csrc/head.cu -- so that the collective choreography of ffc_b200.dist can be
run on CPU under gloo.  It is never imported by the product.
"""
import math

import torch

from oracle.lru_ref import LRU


class _LruShim:
    def __init__(self, cap):
        self.ref = LRU(cap)

    def restore_arrays(self, keys, slots):
        self.ref.restore(list(zip(keys.tolist(), slots.tolist())))

    def state_dict(self):
        return self.ref.state_dict()


class CpuShardBackend:
    dtype = torch.float64

    def __init__(self, feat_dim, q_local, q_total, col_offset, max_rows, scale, loss_type, margin, topk, queue=None):
        self.D, self.Ql, self.off, self.k = feat_dim, q_local, col_offset, topk
        self.scale, self.loss_type, self.margin = scale, loss_type, margin
        self.lru = _LruShim(q_local)
        self.qpos = [0] * q_local
        self.queue = torch.zeros(2, q_local, feat_dim, dtype=torch.float64) if queue is None else queue.double().clone()
        self.ones = []

    def new_stats(self, n, n_ranks):
        return dict(red=torch.zeros(8, n, dtype=torch.float64), osum=torch.zeros(4, n, self.D, dtype=torch.float64),
                    topv=torch.full((n_ranks, 3, n, self.k), -math.inf, dtype=torch.float64),
                    topi=torch.full((n_ranks, 3, n, self.k), -1, dtype=torch.int32))

    # ffc.py:162-177 / 214-235 over the shard's own keys
    def assign(self, keys_compact, n_dev, journal):
        n = int(n_dev.item())
        ref = self.lru.ref
        self.rows, self.cols, self.ones, self.saved = [], [], [], {}
        for k in keys_compact[:n].tolist():
            known = k in ref
            s = ref.try_get(k) if journal else ref.get(k)
            if journal and s not in self.saved:
                self.saved[s] = self.qpos[s]
            if known:
                self.rows.append(self.qpos[s])
                if s not in self.ones:
                    self.ones.append(s)
                self.qpos[s] ^= 1
            else:
                self.rows.append(0)
                self.qpos[s] = 1
            self.cols.append(s)
        self.journal = journal

    def scatter(self, g_all, order, save_undo):
        g_compact = g_all[order]
        last = {}
        for i, rc in enumerate(zip(self.rows, self.cols)):
            last[rc] = i
        self.undo = {rc: self.queue[rc[0], rc[1]].clone() for rc in last} if save_undo else None
        for rc, i in last.items():
            self.queue[rc[0], rc[1]] = g_compact[i].double()

    def undo_bookkeeping(self):
        for s, v in self.saved.items():
            self.qpos[s] = v
        self.lru.ref.rollback_steps(len(self.cols))

    def restore_queue(self):
        for rc, v in self.undo.items():
            self.queue[rc[0], rc[1]] = v

    def view(self, keys):
        return torch.tensor([self.lru.ref.view(k) for k in keys.tolist()], dtype=torch.int32)

    def _row_info(self, label):
        tcol = torch.where((label >= self.off) & (label < self.off + self.Ql), label - self.off, torch.full_like(label, -1)).long()
        pos = {s: j for j, s in enumerate(self.ones)}
        tpos = torch.tensor([pos.get(int(t), -1) for t in tcol.tolist()], dtype=torch.long)
        return tcol, tpos

    SV_T = 1.2        # ffc.py:47 mask_svfc

    def _fixed_max(self):
        """csrc/head.cu fixed_max_of: the largest scaled logit any column can reach (SV lifts hard examples to T*cos + T - 1)"""
        return self.scale * (2 * self.SV_T - 1) if self.loss_type == 'SV' else self.scale

    def prep(self, p_all, label, st, slot):
        """target cosines under queue[0] / W2 and the owner flag (ffc_head_prep)"""
        p = p_all.double()
        n = p.shape[0]
        tcol, tpos = self._row_info(label)
        W0, W1 = self.queue[0], self.queue[1]
        ar = torch.arange(n)
        has_t = tcol >= 0
        t0 = torch.zeros(n, dtype=torch.float64)
        t1 = torch.zeros(n, dtype=torch.float64)
        rows = ar[has_t]
        t0[rows] = (p[rows] * W0[tcol[rows]]).sum(1)
        wt = torch.where((tpos[rows] >= 0).unsqueeze(1), W1[tcol[rows]], W0[tcol[rows]])
        t1[rows] = (p[rows] * wt).sum(1)
        red = st['red']
        red[4], red[5], red[6], red[7] = t0, t1, has_t.double(), torch.zeros(n, dtype=torch.float64)

    def sweep(self, p_all, label, st, slot):
        self.prep(p_all, label, st, slot)
        self.sweep_prepared(p_all, label, st, slot)

    def sweep_prepared(self, p_all, label, st, slot):
        """the sweeps proper; SV reads the (all-reduced) target cosines of st['red'][4:6] for its thresholds (ffc_head_sweep_prepared)"""
        p = p_all.double()
        n, k, s, M = p.shape[0], self.k, self.scale, self._fixed_max()
        sv, T = self.loss_type == 'SV', self.SV_T
        tcol, tpos = self._row_info(label)
        out = label < 0
        W0, W1 = self.queue[0], self.queue[1]
        ones = torch.tensor(self.ones, dtype=torch.long)
        red, osum = st['red'], st['osum']
        cos = p @ W0.t()
        excl = torch.zeros(n, self.Ql, dtype=torch.bool)
        excl[:, ones] = True
        ar = torch.arange(n)
        has_t = tcol >= 0
        excl[ar[has_t], tcol[has_t]] = True
        if sv:   # ffc.py:121-125: columns with cos > gt - margin become T*cos + T - 1 (gradient factor T); thresholds per loss
            inf = torch.full((n,), math.inf, dtype=torch.float64)
            thr = [torch.where(red[6] > 0, red[4 + l] - self.margin, inf) for l in range(2)]

            def terms(c, l):
                hard = c > thr[l].unsqueeze(1)
                z = torch.where(hard, T * c + T - 1.0, c)
                return torch.exp(s * z - M), torch.where(hard, torch.full_like(c, T), torch.ones_like(c))
            for l in range(2):
                pt, coef = terms(cos, l)
                pt = pt.masked_fill(excl, 0.0)
                red[l] = pt.sum(1)
                osum[l] = (pt * coef) @ W0
        else:
            def terms(c, l):
                return torch.exp(s * c - M), torch.ones_like(c)
            pt = torch.exp(s * cos - M).masked_fill(excl, 0.0)
            red[0] = red[1] = pt.sum(1)
            osum[0] = osum[1] = pt @ W0
        st['topv'][slot].fill_(-math.inf)
        st['topi'][slot].fill_(-1)

        def put_topk(vals, idx_map, excl_mask, dst):
            v = vals.masked_fill(excl_mask, -math.inf)
            kk = min(k, v.shape[1])
            if kk == 0:
                return
            tv, ti = torch.topk(v, kk, dim=1)
            gi = idx_map[ti].to(torch.int32)
            gi = torch.where(torch.isinf(tv), torch.full_like(gi, -1), gi)
            st['topv'][slot, dst, out, :kk] = tv[out]
            st['topi'][slot, dst, out, :kk] = gi[out]

        cmask_only = torch.zeros(n, self.Ql, dtype=torch.bool)
        cmask_only[:, ones] = True
        put_topk(cos, self.off + torch.arange(self.Ql), cmask_only, 0)
        for l, W in enumerate((W0, W1)):
            Ws = W[ones]
            cs = p @ Ws.t()
            ex = torch.zeros(n, len(self.ones), dtype=torch.bool)
            has_p = tpos >= 0
            ex[ar[has_p], tpos[has_p]] = True
            ps, coef = terms(cs, l)
            ps = ps.masked_fill(ex, 0.0)
            red[2 + l] = ps.sum(1)
            osum[2 + l] = (ps * coef) @ Ws
            put_topk(cs, self.off + ones, torch.zeros_like(ex), 1 + l)

    def finalize(self, p_all, label, st, n_ranks):
        n, D, k, s, M, m = p_all.shape[0], self.D, self.k, self.scale, self._fixed_max(), self.margin
        tcol, tpos = self._row_info(label)
        out = label < 0
        n_pos, n_out = int((~out).sum()), int(out.sum())
        red, osum = st['red'], st['osum']
        W = self.queue
        loss = torch.zeros((), dtype=torch.float64)
        dp = torch.zeros(n, D, dtype=torch.float64)
        for i in range(n):
            if not out[i]:
                for l in range(2):
                    ct = float(red[4 + l, i])
                    cs = l if self.loss_type == 'SV' else 0       # SV: each loss has its own common statistics (own threshold)
                    if self.loss_type == 'AM':
                        ft, dft = ct - m, 1.0
                    elif self.loss_type == 'SV':
                        ft, dft = (ct - m if ct > m else ct), 1.0  # ffc.py:123 final_gt
                    else:
                        sn = math.sqrt(1.0 - ct * ct)
                        ft, dft = ct * math.cos(m) - sn * math.sin(m), math.cos(m) + ct * math.sin(m) / sn
                    zt = s * ft
                    et = math.exp(zt - M)
                    L = float(red[cs, i] + red[2 + l, i]) + et
                    loss += (math.log(L) + M - zt) / n_pos
                    dp[i] += (s / L / n_pos) * (osum[cs, i] + osum[2 + l, i])
                    if tcol[i] >= 0:
                        row = 1 if (l == 1 and tpos[i] >= 0) else 0
                        dp[i] += (s * (et / L - 1.0) * dft / n_pos) * W[row, tcol[i]]
            else:
                wneg = 1.0 / (n_out * k)
                for l in range(2):
                    cand = []
                    for r in range(n_ranks):
                        for src, setid in ((0, 0), (1, 1 + l)):
                            for q in range(k):
                                v, gi = float(st['topv'][r, setid, i, q]), int(st['topi'][r, setid, i, q])
                                if gi >= 0:
                                    cand.append((v, gi, src))
                    cand.sort(key=lambda t: -t[0])
                    for v, gi, src in cand[:k]:
                        loss += max(v, 0.0) * wneg
                        loc = gi - self.off
                        if v >= 0 and 0 <= loc < self.Ql:
                            dp[i] += wneg * W[l if src else 0, loc]
        return loss.reshape(()), dp

    def end_pass(self):
        pass

    def export_state(self):
        return dict(lru=self.lru.ref.state_dict(), queue=self.queue.clone(), qpos=list(self.qpos))

    def import_state(self, st):
        self.lru = _LruShim(self.Ql)
        self.lru.ref.restore([(int(k), int(v)) for k, v in st['lru']])
        self.queue = st['queue'].double().clone()
        self.qpos = [int(v) for v in st['qpos']]


class _OverlayQueue:
    """queue[row, loc] with substitutions (the overlay of the merged step: csrc/head.cu FinalizeArgs::ovl_map)"""

    def __init__(self, queue, overlay, counter):
        self.queue, self.overlay, self.counter = queue, overlay, counter

    def __getitem__(self, idx):
        key = (int(idx[0]), int(idx[1]))
        v = self.overlay.get(key)
        if v is None:
            return self.queue[key[0], key[1]]
        self.counter[0] += 1
        return v


class CpuShardBackendM(CpuShardBackend):
    """The stand-in speaking the record / merged-step protocol of CudaShardBackend (AM / Arc): three bookkeeping sets, one record
    per pass, finalize from the gathered records, the rollback pass's finalize through an overlay of the rows that restore / the
    commit enqueue have rewritten.  `use_overlay=False` reads the rewritten queue instead (the test shows that this is wrong)."""
    record_path = True
    merged = True
    _SET_ATTRS = ('rows', 'cols', 'ones', 'saved', 'undo', 'journal', 'written')

    def __init__(self, *a, use_overlay=True, **k):
        super().__init__(*a, **k)
        self._sets, self._cur = [dict(), dict(), dict()], 0
        self._st = {}
        self.use_overlay = use_overlay
        self.overlay_reads = 0

    def use_set(self, i):
        self._sets[self._cur] = {k: self.__dict__.get(k) for k in self._SET_ATTRS}
        for k in self._SET_ATTRS:
            self.__dict__[k] = self._sets[i].get(k)
        self._cur = i

    def _of_set(self, i, name):
        return self.__dict__.get(name) if i == self._cur else self._sets[i].get(name)

    def scatter(self, g_all, order, save_undo, overlay_table=None):
        g_compact = g_all[order]
        last = {}
        for i, rc in enumerate(zip(self.rows, self.cols)):
            last[rc] = i
        self.undo = {rc: self.queue[rc[0], rc[1]].clone() for rc in last}        # previous content (restore; the overlay's table 1)
        self.written = {rc: g_compact[i].double().clone() for rc, i in last.items()}   # the overlay's table 0
        for rc, v in self.written.items():
            self.queue[rc[0], rc[1]] = v

    def overlay_clear(self, set_idx):
        pass

    def new_records(self, n, n_ranks, passes=1):
        w = 8 * n + 6 * n * self.k
        return dict(own=torch.zeros(passes, w, dtype=torch.float64), all=torch.zeros(n_ranks, passes, w, dtype=torch.float64), words=w, passes=passes)

    def sweep_record(self, p_all, label, rec, which=0):
        n = p_all.shape[0]
        st = self.new_stats(n, 1)
        self.sweep(p_all, label, st, 0)
        self._st[which] = (st, self._cur)
        rec['own'][which] = torch.cat([st['red'].flatten(), st['topv'][0].flatten(), st['topi'][0].double().flatten()])

    def finalize_gathered(self, p_all, label, rec, n_ranks, which=0, overlay_g=None, route=None):
        n, k = p_all.shape[0], self.k
        allr = rec['all'][:, which]                                        # [R, words]
        own_st, set_idx = self._st[which]
        st = dict(red=allr[:, :8 * n].sum(0).view(8, n), osum=own_st['osum'],
                  topv=allr[:, 8 * n:8 * n + 3 * n * k].reshape(n_ranks, 3, n, k),
                  topi=allr[:, 8 * n + 3 * n * k:].reshape(n_ranks, 3, n, k).to(torch.int32))
        real = self.queue
        if overlay_g is not None and self.use_overlay:
            # rows the commit enqueue (set 1) replaced: their previous content ... unless the rollback enqueue had its own row there
            overlay = dict(self._of_set(1, 'undo') or {})
            overlay.update(self._of_set(set_idx, 'written') or {})
            hits = [0]
            self.queue = _OverlayQueue(real, overlay, hits)
        try:
            # finalize reads the bookkeeping of ITS pass (`ones` positions of the target columns)
            cur = self._cur
            self.use_set(set_idx)
            out = self.finalize(p_all, label, st, n_ranks)
            self.use_set(cur)
            return out
        finally:
            if self.queue is not real:
                self.overlay_reads += self.queue.counter[0]
                self.queue = real
