"""Optional mode `queue_grad=True` (north_star: "dQueue where the queue is trainable"; the reference's queue is a no-grad buffer,
ffc.py:29): d(loss of a pass)/d(queue) from the swapped tcgen05 sweep + the special-row / hard-negative kernels, against autograd
on a queue that requires grad in the oracle (oracle/head_ref.py dqueue_ref, fp64).  Off by default: nothing changes for the parity
path."""
import pytest
import torch
import torch.nn.functional as F

from oracle.head_ref import HeadOracle, dqueue_ref

pytestmark = pytest.mark.gpu


def _rel(a, b):
    return float((a - b).norm() / (b.norm() + 1e-30))


def _center(i, D, seed):
    return torch.randn(D, generator=torch.Generator().manual_seed(seed * 1000003 + int(i)))


def _clustered(ids, D, seed, noise, gen):
    c = torch.stack([_center(i, D, seed) for i in ids.tolist()])
    return F.normalize(F.normalize(c) + noise * torch.randn(len(ids), D, generator=gen) / D ** 0.5)


@pytest.mark.parametrize('D,Q,B,n_ids,loss_type,margin', [(128, 1000, 96, 700, 'Arc', 0.5), (512, 4096, 256, 3000, 'AM', 0.4), (256, 2000, 128, 5000, 'Arc', 0.5)])
def test_dqueue_matches_autograd_on_a_trainable_queue(D, Q, B, n_ids, loss_type, margin):
    """commit passes (the queue a commit pass leaves behind is the queue it swept): targets, `ones` slots (hits), in-batch duplicate
    labels and -- in the third configuration, identities > queue -- outlier rows with their hard negatives"""
    import ffc_b200
    dev = torch.device('cuda')
    torch.manual_seed(0)
    h = ffc_b200.FFCHead(D, Q, 32.0, loss_type, margin, precision='bf16', max_batch=B, device=dev, queue_grad=True)
    h._ensure()
    o = HeadOracle(D, Q, 32.0, loss_type, margin, queue=h.queue.cpu(), dtype=torch.float64)
    gen = torch.Generator().manual_seed(4)
    acc = torch.zeros(2, Q, D, dtype=torch.float64)
    for step in range(4):
        xl = torch.randint(0, n_ids, (B,), generator=gen)
        yl = torch.cat([xl[:B // 2], torch.randint(0, n_ids, (B - B // 2,), generator=gen)])
        x = _clustered(xl, D, 5, 0.7, gen)
        y = _clustered(yl, D, 5, 0.7, gen)
        loss = h.head(x.to(dev), y.to(dev), xl, yl, commit=True)
        lo = o.head_pass(x.double(), y.double(), xl.tolist(), yl.tolist(), commit=True)
        tr = o.trace[-1]
        assert h.last_bookkeeping()[2] == tr['labels']
        l_ref, dq_ref = dqueue_ref(x.double(), o.queue, tr['labels'], tr['ones'], loss_type, margin, 32.0, o.k)
        assert abs(float(l_ref) - float(lo)) <= 1e-9 * abs(float(lo))
        assert abs(float(loss) - float(lo)) <= 1e-2 * abs(float(lo))
        got = h.dqueue_pass.double().cpu()
        assert torch.isfinite(got).all()
        # every region on its own: untouched rows (the swapped sweep), target / `ones` rows (the special-row kernel), queue[1]
        special = sorted(set(tr['ones']) | {l for l in tr['labels'] if l >= 0})
        plain = torch.ones(Q, dtype=torch.bool)
        plain[special] = False
        n_out = sum(l < 0 for l in tr['labels'])
        if n_out == 0:
            assert _rel(got[0][plain], dq_ref[0][plain]) <= 1e-2, ('bulk rows', step, _rel(got[0][plain], dq_ref[0][plain]))
            assert _rel(got[0][special], dq_ref[0][special]) <= 1e-2, ('special rows, queue[0]', step)
            assert _rel(got[1], dq_ref[1]) <= 1e-2, ('queue[1]', step)
        # with outliers a near-tie at a row's top-k boundary moves a whole prototype row between two slots: compare in aggregate
        assert _rel(got, dq_ref) <= (1e-2 if n_out == 0 else 5e-2), (step, _rel(got, dq_ref), n_out)
        acc += dq_ref
    assert _rel(h.dqueue.double().cpu(), acc) <= 3e-2            # accumulated over the passes
    h.zero_dqueue()
    assert float(h.dqueue.abs().max()) == 0.0


def test_dqueue_is_off_by_default():
    import ffc_b200
    h = ffc_b200.FFCHead(64, 256, 32.0, 'AM', 0.4, precision='bf16', max_batch=32, device=torch.device('cuda'))
    h._ensure()
    assert h.dqueue is None and h.queue_grad is False
    with pytest.raises(ffc_b200.FFCError):
        from ffc_b200 import _capi
        import ctypes as C
        hp = _capi.HeadPass()
        _capi.check(_capi.lib().ffc_head_dqueue(h._h, C.byref(hp), h.queue.data_ptr(), None))
