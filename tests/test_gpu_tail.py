"""SURVEY 8(f) rank 3 on the GPU: the backbone tail (BatchNorm1d + F.normalize, forward and backward) through the C ABI
(`ffc_tail_forward` / `ffc_tail_backward`) against the oracle (oracle/tail_ref.py, float64) and against the fixtures captured
from the reference's own backbones (tests/golden/tail_*.npz).  Floating point: 2e-5 of the tensor's max-abs (fp32 kernels)."""
import glob
import os

import numpy as np
import pytest
import torch
import torch.nn as nn
import torch.nn.functional as F

from oracle import tail_ref

pytestmark = pytest.mark.gpu

TOL = 2e-5
CASES = sorted(os.path.basename(p)[5:-4] for p in glob.glob(os.path.join(os.path.dirname(__file__), 'golden', 'tail_*.npz')))


def close(a, b, what, tol=TOL):
    a = a.detach().double().cpu().numpy() if torch.is_tensor(a) else np.asarray(a, dtype=np.float64)
    b = b.detach().double().cpu().numpy() if torch.is_tensor(b) else np.asarray(b, dtype=np.float64)
    assert a.shape == b.shape, (what, a.shape, b.shape)
    err = np.abs(a - b).max()
    assert err <= tol * max(np.abs(b).max(), 1e-30), (what, err, np.abs(b).max())


def dev_t(a):
    return torch.tensor(np.asarray(a), dtype=torch.float32, device='cuda')


@pytest.mark.parametrize('case', CASES)
def test_tail_matches_reference_backbone_fixture(case, golden_dir):
    import ffc_b200
    z = np.load(os.path.join(golden_dir, f'tail_{case}.npz'))
    if not bool(z['bn']):
        x = dev_t(z['x0']).requires_grad_(True)
        p = ffc_b200.l2_normalize(x)
        p.backward(dev_t(z['dp0']))
        close(p, z['p0'], 'p')
        close(x.grad, z['dx0'], 'dx')
        return
    D = z['weight'].shape[0]
    tail = ffc_b200.FFCTail(D, eps=float(z['eps']), momentum=float(z['momentum'])).cuda()
    with torch.no_grad():
        tail.weight.copy_(dev_t(z['weight']))
        tail.bias.copy_(dev_t(z['bias']))
    tail.weight.requires_grad = False                          # resnet_arcface.py:101
    tail.train(bool(z['training']))
    for s in range(int(z['steps'])):
        with torch.no_grad():
            tail.running_mean.copy_(dev_t(z[f'rm_in{s}']))
            tail.running_var.copy_(dev_t(z[f'rv_in{s}']))
        x = dev_t(z[f'x{s}']).requires_grad_(True)
        tail.zero_grad()
        p = tail(x)
        p.backward(dev_t(z[f'dp{s}']))
        close(p, z[f'p{s}'], 'p')
        close(x.grad, z[f'dx{s}'], 'dx')
        close(tail.bias.grad, z[f'dbias{s}'], 'dbias')
        close(tail.running_mean, z[f'rm_out{s}'], 'running_mean')
        close(tail.running_var, z[f'rv_out{s}'], 'running_var')
        assert tail.weight.grad is None
    assert int(tail.num_batches_tracked) == (int(z['steps']) if bool(z['training']) else 0)


@pytest.mark.parametrize('B,D', [(8, 64), (37, 130), (513, 256), (1024, 512), (1500, 128), (5000, 64)])
@pytest.mark.parametrize('mode', ['normalize', 'bn_train', 'bn_eval'])
def test_tail_matches_oracle(B, D, mode):
    """seeded inputs, every mode; D = 130 takes the scalar (unvectorised) kernels; weight and bias gradients included"""
    import ffc_b200
    rng = np.random.default_rng(B * 1000 + D)
    x_np = rng.normal(0.3, 1.7, size=(B, D)).astype(np.float32)
    dp_np = rng.normal(size=(B, D)).astype(np.float32)
    x = dev_t(x_np).requires_grad_(True)
    if mode == 'normalize':
        p = ffc_b200.NormalizeTail()(x)
        p.backward(dev_t(dp_np))
        p_ref, cache, _, _ = tail_ref.tail_forward(x_np, bn=False)
        dx_ref, _, _ = tail_ref.tail_backward(dp_np, cache)
        close(p, p_ref, 'p')
        close(x.grad, dx_ref, 'dx')
        return
    w = rng.uniform(0.5, 1.5, D).astype(np.float32)
    b = rng.normal(0, 0.3, D).astype(np.float32)
    rm = rng.normal(0, 0.2, D).astype(np.float32)
    rv = rng.uniform(0.5, 1.5, D).astype(np.float32)
    tail = ffc_b200.FFCTail(D).cuda()
    with torch.no_grad():
        tail.weight.copy_(dev_t(w))
        tail.bias.copy_(dev_t(b))
        tail.running_mean.copy_(dev_t(rm))
        tail.running_var.copy_(dev_t(rv))
    tail.train(mode == 'bn_train')
    p = tail(x)
    p.backward(dev_t(dp_np))
    p_ref, cache, rm_ref, rv_ref = tail_ref.tail_forward(x_np, w, b, rm, rv, training=(mode == 'bn_train'))
    dx_ref, dw_ref, db_ref = tail_ref.tail_backward(dp_np, cache)
    close(p, p_ref, 'p')
    close(x.grad, dx_ref, 'dx')
    close(tail.weight.grad, dw_ref, 'dweight')
    close(tail.bias.grad, db_ref, 'dbias')
    close(tail.running_mean, rm_ref, 'running_mean')
    close(tail.running_var, rv_ref, 'running_var')


def test_tail_writes_into_a_packed_staging_buffer():
    """out=: rows land in a strided view (the sharded head's all-gather input is [x | y | labels] in one buffer)"""
    import ffc_b200
    B, D = 64, 128
    torch.manual_seed(3)
    x = torch.randn(B, D, device='cuda', requires_grad=True)
    y = torch.randn(B, D, device='cuda', requires_grad=True)
    tail = ffc_b200.FFCTail(D).cuda()
    stage = torch.full((B, 2 * D + 8), 7.0, device='cuda')
    px = tail(x, out=stage[:, :D])
    py = tail(y, out=stage[:, D:2 * D])
    assert px.data_ptr() == stage.data_ptr() and py.data_ptr() == stage.data_ptr() + 4 * D
    ref = nn.BatchNorm1d(D).cuda()
    close(stage[:, :D], F.normalize(ref(x)), 'x half')
    close(stage[:, D:2 * D], F.normalize(ref(y)), 'y half')
    assert bool((stage[:, 2 * D:] == 7.0).all())
    dp = torch.randn(B, 2 * D, device='cuda')
    (px * dp[:, :D]).sum().backward()                         # strided dp as well
    xr = x.detach().clone().requires_grad_(True)
    ref2 = nn.BatchNorm1d(D).cuda()
    (F.normalize(ref2(xr)) * dp[:, :D]).sum().backward()
    close(x.grad, xr.grad, 'dx')


def test_tail_properties_at_full_size():
    """B = 8192 rows (8 ranks x 1024), D = 512: unit norms, dy orthogonal to p, BatchNorm's two batch constraints on dx"""
    import ffc_b200
    B, D = 8192, 512
    torch.manual_seed(5)
    x = (torch.randn(B, D, device='cuda') * 3 + 1).requires_grad_(True)
    dp = torch.randn(B, D, device='cuda')
    p = ffc_b200.l2_normalize(x)
    p.backward(dp)
    assert float((p.double().norm(dim=1) - 1).abs().max()) < 1e-6
    assert float(((x.grad.double() * p.double()).sum(1)).abs().max()) < 1e-5 * float(x.grad.abs().max()) * D ** 0.5
    tail = ffc_b200.FFCTail(D).cuda()
    x2 = x.detach().clone().requires_grad_(True)
    p2 = tail(x2)
    p2.backward(dp)
    assert float((p2.double().norm(dim=1) - 1).abs().max()) < 1e-6
    g = x2.grad.double()
    xhat = (x2.detach().double() - x2.detach().double().mean(0)) / x2.detach().double().var(0, unbiased=False).add(tail.eps).sqrt()
    scale = float(g.abs().max()) * B
    assert float(g.sum(0).abs().max()) < 1e-5 * scale                     # sum_b dx = 0
    assert float((g * xhat).sum(0).abs().max()) < 1e-5 * scale            # sum_b dx * xhat = 0
    # idempotent on unit-norm input: what the head's callers rely on when they keep their own F.normalize
    close(ffc_b200.l2_normalize(p.detach()), p, 'idempotent', tol=1e-6)


def test_tail_is_run_to_run_bit_identical():
    """every reduction runs in a fixed order (slab partials combined in slab order, no floating-point atomics)"""
    import ffc_b200
    B, D = 3000, 512
    torch.manual_seed(8)
    x0 = torch.randn(B, D, device='cuda') * 2 + 0.5
    dp = torch.randn(B, D, device='cuda')
    runs = []
    for _ in range(3):
        tail = ffc_b200.FFCTail(D).cuda()
        x = x0.clone().requires_grad_(True)
        p = tail(x)
        p.backward(dp)
        runs.append((p.detach().clone(), x.grad.clone(), tail.weight.grad.clone(), tail.bias.grad.clone(), tail.running_var.clone()))
        torch.empty(64 << 20, device='cuda').zero_()          # perturb the schedule between runs
    for r in runs[1:]:
        for a, b in zip(runs[0], r):
            assert torch.equal(a, b)


def test_tail_clamped_row_and_errors():
    import ffc_b200
    x = torch.zeros(2, 4, device='cuda')
    x[1] = torch.tensor([3., 0, 4, 0])
    x.requires_grad_(True)
    p = ffc_b200.l2_normalize(x)
    p.sum().backward()
    xr = x.detach().clone().requires_grad_(True)
    F.normalize(xr).sum().backward()
    assert torch.equal(p[0], torch.zeros(4, device='cuda'))
    close(x.grad, xr.grad, 'dx with a clamped row')               # row 0: dp / 1e-12
    tail = ffc_b200.FFCTail(4).cuda()
    with pytest.raises(ValueError, match='more than 1 value per channel'):
        tail(torch.randn(1, 4, device='cuda'))
    tail.eval()
    tail(torch.randn(1, 4, device='cuda'))                         # eval mode takes a single row, like nn.BatchNorm1d
    with pytest.raises(ValueError):
        tail(torch.randn(3, 5, device='cuda'))
    with pytest.raises(ffc_b200.FFCError):
        ffc_b200.l2_normalize(torch.randn(3, 4))                   # CPU tensor: no fallback


def test_tail_is_a_batchnorm1d_drop_in():
    """same state_dict as nn.BatchNorm1d (reference checkpoints load), shared parameters through fuse_tail, momentum=None,
    track_running_stats=False, half-precision input, launch counts"""
    import ffc_b200
    from ffc_b200 import _capi
    D, B = 64, 48
    torch.manual_seed(9)

    class Net(nn.Module):                                           # the shape of resnet_arcface.py:98-101,150-151
        def __init__(self):
            super().__init__()
            self.fc = nn.Linear(32, D)
            self.features = nn.BatchNorm1d(D, eps=1e-05)
            nn.init.constant_(self.features.weight, 1.0)
            self.features.weight.requires_grad = False

        def forward(self, x):
            return F.normalize(self.features(self.fc(x)))

    ref, fused = Net().cuda(), Net().cuda()
    fused.load_state_dict(ref.state_dict())
    bias_before = fused.features.bias
    ffc_b200.fuse_tail(fused)
    assert isinstance(fused.features, ffc_b200.FFCTail) and fused.features.bias is bias_before
    assert list(fused.state_dict().keys()) == list(ref.state_dict().keys())
    assert not fused.features.weight.requires_grad
    for step in range(3):
        x = torch.randn(B, 32, device='cuda')
        dp = torch.randn(B, D, device='cuda')
        for net in (ref, fused):
            net.zero_grad()
            (net(x) * dp).sum().backward()
        close(fused.fc.weight.grad, ref.fc.weight.grad, 'fc.weight.grad', tol=1e-4)
        close(fused.features.bias.grad, ref.features.bias.grad, 'features.bias.grad', tol=1e-4)
    for k, v in ref.state_dict().items():
        close(fused.state_dict()[k].float(), v.float(), k)
    ref.eval(), fused.eval()
    x = torch.randn(5, 32, device='cuda')
    close(fused(x), ref(x), 'eval forward')

    for kw in (dict(momentum=None), dict(track_running_stats=False), dict(affine=False)):
        a, b = nn.BatchNorm1d(D, **kw).cuda(), ffc_b200.FFCTail(D, **kw).cuda()
        for step in range(3):
            x = torch.randn(B, D, device='cuda')
            close(b(x), F.normalize(a(x)), str(kw))
        for k, v in a.state_dict().items():
            close(b.state_dict()[k].float(), v.float(), f'{kw} {k}')
        a.eval(), b.eval()
        x = torch.randn(7, D, device='cuda')
        close(b(x), F.normalize(a(x)), f'{kw} eval')

    xh = torch.randn(B, D, device='cuda', dtype=torch.float16, requires_grad=True)      # autocast-style half input
    t = ffc_b200.FFCTail(D).cuda()
    out = t(xh)
    out.sum().backward()
    assert out.dtype == torch.float32 and xh.grad.dtype == torch.float16

    lib = _capi.lib()
    x = torch.randn(B, D, device='cuda', requires_grad=True)
    n0 = lib.ffc_launch_count()
    p = t(x)
    n1 = lib.ffc_launch_count()
    p.sum().backward()
    n2 = lib.ffc_launch_count()
    assert (n1 - n0, n2 - n1) == (2, 2)
    n0 = lib.ffc_launch_count()
    p = ffc_b200.l2_normalize(x)
    n1 = lib.ffc_launch_count()
    p.sum().backward()
    assert (n1 - n0, lib.ffc_launch_count() - n1) == (1, 1)


def test_tail_feeds_the_head():
    """probe / gallery nets ending in FFCTail through FFC.forward: same loss and probe gradients as the eager tail"""
    import ffc_b200
    D, B, Q = 64, 32, 256
    torch.manual_seed(21)

    class Net(nn.Module):
        def __init__(self, fused):
            super().__init__()
            self.fc = nn.Linear(40, D)
            self.features = ffc_b200.FFCTail(D) if fused else nn.BatchNorm1d(D)
            self.fused = fused

        def forward(self, x):
            x = self.features(self.fc(x))
            return x if self.fused else F.normalize(x)

    out = []
    sd = Net(False).state_dict()
    for fused in (False, True):
        torch.manual_seed(22)
        nets = Net(fused), Net(fused)
        for n in nets:
            n.load_state_dict(sd)
        m = ffc_b200.FFC('x', D, queue_size=Q, loss_type='AM', margin=0.4, probe_net=nets[0], gallery_net=nets[1], max_batch=B,
                         precision='fp32').cuda()
        g = torch.Generator().manual_seed(23)
        losses = []
        for step in range(2):
            x, y = torch.randn(B, 40, generator=g).cuda(), torch.randn(B, 40, generator=g).cuda()
            xl, yl = torch.randint(0, 400, (B,), generator=g), torch.randint(0, 400, (B,), generator=g)
            m.zero_grad()
            loss = m(x, y, xl, yl)
            loss.backward()
            losses.append(float(loss))
        out.append((losses, m.probe_net.fc.weight.grad.clone(), m.probe_net.features.bias.grad.clone(), m.queue.clone()))
    (l0, gw0, gb0, q0), (l1, gw1, gb1, q1) = out
    assert np.allclose(l0, l1, rtol=1e-5), (l0, l1)
    close(gw1, gw0, 'fc.weight.grad', tol=1e-4)
    close(gb1, gb0, 'features.bias.grad', tol=1e-4)
    close(q1, q0, 'queue', tol=1e-5)
