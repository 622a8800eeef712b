"""SURVEY 8(f) ranks 1 and 2 on the GPU: the one-launch gallery EMA is bit-identical to the reference expression
(ffc.py:139-145), and the reference's checkpoint dict (main.py:84-85) round-trips through FFC.load_checkpoint."""
import io

import pytest
import torch
import torch.nn as nn
import torch.nn.functional as F

pytestmark = pytest.mark.gpu


class _Net(nn.Module):
    """a backbone-shaped bag of parameters (odd sizes, conv / bn / linear) ending in the L2 normalisation every reference net has"""

    def __init__(self, D):
        super().__init__()
        self.c1 = nn.Conv2d(3, 17, 3)
        self.bn = nn.BatchNorm2d(17)
        self.fc = nn.Linear(D, D)
        self.extra = nn.Parameter(torch.randn(70001))      # > one EMA chunk, not a multiple of 4

    def forward(self, x):
        return F.normalize(self.fc(x))


def test_gallery_ema_bit_exact():
    import ffc_b200
    dev = torch.device('cuda')
    torch.manual_seed(0)
    D = 64
    m = ffc_b200.FFC('x', D, queue_size=128, momentum=0.99, probe_net=_Net(D), gallery_net=_Net(D), max_batch=16).to(dev)
    ref = [p.detach().clone() for p in m.gallery_net.parameters()]
    for step in range(3):
        with torch.no_grad():
            for p in m.probe_net.parameters():
                p.add_(0.01 * torch.randn_like(p))
        m._momentum_update_gallery()
        for i, (pp, pg) in enumerate(zip(m.probe_net.parameters(), ref)):
            ref[i] = pg * m.m + pp.data * (1. - m.m)                      # ffc.py:145, verbatim
        for got, want in zip(m.gallery_net.parameters(), ref):
            assert torch.equal(got.data, want), step
    assert all(not p.requires_grad for p in m.gallery_net.parameters())


def test_checkpoint_round_trip_resumes_identically():
    import ffc_b200
    dev = torch.device('cuda')
    D, Q, B, n_ids = 64, 256, 32, 400

    def make():
        torch.manual_seed(1)
        return ffc_b200.FFC('identity', D, queue_size=Q, loss_type='Arc', margin=0.5, precision='fp32', max_batch=B).to(dev)

    a = make()
    gen = torch.Generator().manual_seed(2)
    batches = []
    for _ in range(5):
        xl = torch.randint(0, n_ids, (B,), generator=gen)
        yl = torch.cat([xl[:B // 2], torch.randint(0, n_ids, (B - B // 2,), generator=gen)])
        batches.append((torch.randn(B, D, generator=gen), torch.randn(B, D, generator=gen), xl, yl))
    for x, y, xl, yl in batches[:3]:
        a(x.to(dev), y.to(dev), xl, yl)
    buf = io.BytesIO()
    torch.save(a.checkpoint(), buf)                                         # the dict main.py:85 writes
    ck = torch.load(io.BytesIO(buf.getvalue()), weights_only=False)
    assert set(ck) == {'state_dict', 'lru', 'fc', 'qp'} and ck['fc'].shape == (2, Q, D) and len(ck['qp']) == Q
    b = make()
    with torch.no_grad():
        b.queue.add_(1.0)                                                   # make sure the load really overwrites
    b.load_checkpoint(ck)
    assert b.lru.state_dict() == a.lru.state_dict() and torch.equal(a.queue, b.queue) and a.queue_position_dict == b.queue_position_dict
    for x, y, xl, yl in batches[3:]:
        xa, xb = x.to(dev).requires_grad_(True), x.to(dev).requires_grad_(True)
        la = a(xa, y.to(dev), xl, yl)
        lb = b(xb, y.to(dev), xl, yl)
        la.backward()
        lb.backward()
        assert abs(float(la) - float(lb)) <= 1e-6 * abs(float(la))
        assert torch.allclose(xa.grad, xb.grad, rtol=1e-5, atol=1e-8)
        assert b.lru.state_dict() == a.lru.state_dict() and torch.equal(a.queue, b.queue) and a.queue_position_dict == b.queue_position_dict
