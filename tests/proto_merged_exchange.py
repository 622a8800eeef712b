"""TEST-ONLY prototype (not product code): the sharded head with ONE exchange point per step instead of one per pass.

Today (ffc_b200/dist.py) a step is   sweep_rb -> exchange -> finalize_rb -> reduce-scatter -> restore -> sweep_cm -> exchange -> finalize_cm ->
reduce-scatter: the ranks meet five times.  Nothing in the commit pass's sweep depends on the rollback pass's exchange, only on the
queue rows being restored; what stands in the way of   sweep_rb -> restore -> sweep_cm -> exchange(both) -> finalize(both) -> reduce(both)
is that finalize reads queue rows -- each positive row's target prototype and each outlier row's hard negatives (the dLoss/dp terms of
ffc.py:83, 88-90 through autograd) -- as they were DURING the pass's sweep, and by then `restore` and the commit pass's enqueue have
overwritten up to 2 * (own keys) of them.  The fix prototyped here: finalize of the rollback pass reads through an OVERLAY,
    (row, slot) written by the rollback enqueue        -> the enqueued gallery row (still in the gathered batch)
    (row, slot) written by the commit enqueue only     -> the row's previous content, saved by the commit enqueue
    anything else                                       -> the queue,
i.e. <= 2 * own-keys entries per rank.  tests/test_dist_cpu.py::test_merged_exchange_prototype runs it under gloo against the dense
oracle; the CUDA finalize kernels do not have the overlay yet."""
import torch
import torch.distributed as dist

from cpu_shard_backend import CpuShardBackend
from ffc_b200.dist import ShardedFFCHead

_SET_ATTRS = ('rows', 'cols', 'ones', 'saved', 'undo', 'journal', 'written')


class OverlayQueue:
    """queue[row, loc] with substitutions"""

    def __init__(self, queue, overlay):
        self.queue, self.overlay = queue, overlay

    def __getitem__(self, idx):
        row, loc = int(idx[0]), int(idx[1])
        v = self.overlay.get((row, loc))
        return self.queue[row, loc] if v is None else v


class CpuShardBackend2(CpuShardBackend):
    """two bookkeeping sets (like CudaShardBackend) + enqueue that records what it wrote + finalize through an overlay"""

    def __init__(self, *a, **k):
        super().__init__(*a, **k)
        self._sets, self._cur = [dict(), dict()], 0
        self.overlay_reads = 0

    def use_set(self, i):
        self._sets[self._cur] = {k: self.__dict__.get(k) for k in _SET_ATTRS}
        for k, v in self._sets[i].items():
            self.__dict__[k] = v
        self._cur = i

    def scatter(self, g_all, order, save_undo):
        g_compact = g_all[order]
        last = {}
        for i, rc in enumerate(zip(self.rows, self.cols)):
            last[rc] = i
        self.undo = {rc: self.queue[rc[0], rc[1]].clone() for rc in last}        # previous content, always (the overlay needs the commit's too)
        self.written = {rc: g_compact[i].double().clone() for rc, i in last.items()}
        for rc, v in self.written.items():
            self.queue[rc[0], rc[1]] = v

    def finalize(self, p_all, label, st, n_ranks, overlay=None):
        if overlay is None:
            return super().finalize(p_all, label, st, n_ranks)
        real = self.queue
        hits = [0]

        class Counting(OverlayQueue):
            def __getitem__(s, idx):
                if (int(idx[0]), int(idx[1])) in s.overlay:
                    hits[0] += 1
                return super().__getitem__(idx)
        self.queue = Counting(real, overlay)
        try:
            return super().finalize(p_all, label, st, n_ranks)
        finally:
            self.queue = real
            self.overlay_reads += hits[0]


class MergedShardedHead(ShardedFFCHead):
    def __init__(self, *a, use_overlay=True, **k):
        super().__init__(*a, **k)
        self.use_overlay = use_overlay

    def _exchange(self, st, n):
        be, R = self.backend, self.R
        dist.all_reduce(st['red'][:4], group=self.group)
        if self.loss_type != 'SV':
            dist.all_reduce(st['red'][4:], group=self.group)
        if R > 1:
            tv, ti = st['topv'][self.rank].clone(), st['topi'][self.rank].clone()
            dist.all_gather_into_tensor(st['topv'].view(R * 3, n, -1), tv, group=self.group)
            dist.all_gather_into_tensor(st['topi'].view(R * 3, n, -1), ti, group=self.group)

    def _sweep(self, p_all, label, st):
        be = self.backend
        if self.loss_type == 'SV':          # the target cosines are needed before the sweep: this all-reduce stays per pass
            be.prep(p_all, label, st, self.rank)
            dist.all_reduce(st['red'][4:], group=self.group)
            be.sweep_prepared(p_all, label, st, self.rank)
        else:
            be.sweep(p_all, label, st, self.rank)

    def forward_pair(self, x, y, x_label, y_label):
        be, R = self.backend, self.R
        x_all, xl_all, y_all, yl_all = self.gather_pair(x, y, x_label, y_label)
        n = x_all.shape[0]
        ctx_rb = self._bookkeep(xl_all, yl_all, False, 0)
        ctx_cm = self._bookkeep(yl_all, xl_all, True, 1)
        st_rb, st_cm = be.new_stats(n, R), be.new_stats(n, R)
        # rollback pass: enqueue, sweep, restore -- no exchange yet
        be.use_set(0)
        be.scatter(y_all, ctx_rb['order'], save_undo=True)
        self._sweep(x_all, ctx_rb['label'], st_rb)
        written_rb = dict(be.written)
        be.restore_queue()
        # commit pass: enqueue, sweep
        be.use_set(1)
        be.scatter(x_all, ctx_cm['order'], save_undo=True)
        self._sweep(y_all, ctx_cm['label'], st_cm)
        overlay = dict(be.undo)              # rows the commit enqueue replaced: their previous content ...
        overlay.update(written_rb)           # ... unless the rollback enqueue had its own row there during its sweep
        # the single exchange point of the step, then both finalizes and one reduction of both gradients
        self._exchange(st_rb, n)
        self._exchange(st_cm, n)
        be.use_set(0)
        l2, dx_part = be.finalize(x_all, ctx_rb['label'], st_rb, R, overlay=overlay if self.use_overlay else None)
        be.use_set(1)
        l1, dy_part = be.finalize(y_all, ctx_cm['label'], st_cm, R)
        both = torch.stack([dx_part, dy_part])
        dist.all_reduce(both, group=self.group)
        B = n // R
        sl = slice(self.rank * B, (self.rank + 1) * B)
        self._last = dict(label=ctx_cm['label'], n_mine=ctx_cm['n_mine'])
        return l1 + l2, both[0, sl].clone(), both[1, sl].clone()
