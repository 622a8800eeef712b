"""GPU parity of the head over shapes the fixtures do not cover (tile edges, D=64..512, outliers, ones set)."""
import pytest
import torch
import torch.nn.functional as F

from oracle.head_ref import HeadOracle

pytestmark = pytest.mark.gpu


def _rel(a, b):
    return float((a - b).norm() / (b.norm() + 1e-30))


def _batch(gen, cen, B, n_ids, noise):
    h = B // 2
    ids = torch.randperm(n_ids, generator=gen)[:h]
    xl = torch.cat([ids, torch.randint(0, n_ids, (B - h,), generator=gen)])
    yl = torch.cat([ids, torch.randint(0, n_ids, (B - h,), generator=gen)])
    D = cen.shape[1]
    x = F.normalize(cen[xl] + noise * torch.randn(B, D, generator=gen))
    y = F.normalize(cen[yl] + noise * torch.randn(B, D, generator=gen))
    return x, y, xl, yl


def _run(D, Q, B, n_ids, loss_type, margin, precision, tol, steps=3, seed=0, noise=0.8):
    import ffc_b200
    dev = torch.device('cuda')
    torch.manual_seed(seed)
    m = ffc_b200.FFC('identity', D, queue_size=Q, loss_type=loss_type, margin=margin, precision=precision, max_batch=B).to(dev)
    o = HeadOracle(D, Q, 32.0, loss_type, margin, queue=m.queue.cpu(), dtype=torch.float64)
    gen = torch.Generator().manual_seed(seed + 1)
    cen = F.normalize(torch.randn(n_ids, D, generator=gen))
    for s in range(steps):
        x, y, xl, yl = _batch(gen, cen, B, n_ids, noise)
        xd = x.to(dev).requires_grad_(True)
        yd = y.to(dev).requires_grad_(True)
        px, py = F.normalize(xd), F.normalize(yd)
        l2 = m.head(px, py.detach(), xl, yl, commit=False)
        rb = m.last_bookkeeping()
        l1 = m.head(py, px.detach(), yl, xl, commit=True)
        cm = m.last_bookkeeping()
        gx, gy = torch.autograd.grad(l1 + l2, [px, py])
        x64 = x.double().requires_grad_(True)
        y64 = y.double().requires_grad_(True)
        ref = o.forward(x64, y64, xl.tolist(), yl.tolist())
        ref.backward()
        for got, want in ((rb, o.trace[-2]), (cm, o.trace[-1])):
            assert got[0] == want['rows'] and got[1] == want['cols'] and got[2] == want['labels'] and got[3] == want['ones'], s
        assert abs(float(l1 + l2) - float(ref)) <= tol * abs(float(ref)), (s, float(l1 + l2), float(ref))
        assert _rel(gx.double().cpu(), x64.grad) <= tol, (s, 'dx', _rel(gx.double().cpu(), x64.grad))
        assert _rel(gy.double().cpu(), y64.grad) <= tol, (s, 'dy', _rel(gy.double().cpu(), y64.grad))


@pytest.mark.parametrize('D,Q,B,n_ids,loss_type,margin', [
    (128, 512, 128, 2000, 'Arc', 0.5),      # exact tiles, many outliers (n_ids >> Q)
    (128, 1000, 96, 900, 'AM', 0.4),        # ragged rows and columns, mostly hits after warm-up
    (64, 300, 40, 250, 'Arc', 0.5),
    (256, 2048, 200, 3000, 'AM', 0.4),
    (512, 4096, 256, 4096, 'Arc', 0.5),     # the headline feature dim; queue == identity count
    (512, 1500, 130, 5000, 'AM', 0.4),
    (128, 700, 64, 600, 'SV', 0.4),
])
def test_bf16_shapes(D, Q, B, n_ids, loss_type, margin):
    _run(D, Q, B, n_ids, loss_type, margin, 'bf16', 1e-2)


@pytest.mark.parametrize('D,Q,B,n_ids,loss_type,margin', [
    (512, 4096, 256, 4096, 'Arc', 0.5),
    (20, 333, 50, 400, 'SV', 0.4),
    (256, 50000, 64, 60000, 'AM', 0.4),     # hard_neg k = 10
])
def test_check_mode_shapes(D, Q, B, n_ids, loss_type, margin):
    _run(D, Q, B, n_ids, loss_type, margin, 'fp32', 1e-5, steps=2)
