"""GPU parity of the FFC head against the golden fixtures (reference outputs) and the fp64 oracle."""
import glob
import os

import numpy as np
import pytest
import torch
import torch.nn.functional as F

from oracle.head_ref import HeadOracle

pytestmark = pytest.mark.gpu

CASES = sorted(os.path.basename(p)[4:-4] for p in glob.glob(os.path.join(os.path.dirname(__file__), 'golden', 'ffc_*.npz')))


def _rel(a, b):
    return float((a - b).norm() / (b.norm() + 1e-30))


def _run_case(z, precision, tol_loss, tol_grad):
    import ffc_b200
    dev = torch.device('cuda')
    D, Q, B = int(z['D']), int(z['Q']), int(z['B'])
    m = ffc_b200.FFC('identity', D, queue_size=Q, scale=float(z['scale']), loss_type=str(z['loss_type']), margin=float(z['margin']),
                     precision=precision, max_batch=max(B, 16))
    m.queue.copy_(torch.from_numpy(z['queue0']))
    m = m.to(dev)
    o64 = HeadOracle(D, Q, float(z['scale']), str(z['loss_type']), float(z['margin']), queue=torch.from_numpy(z['queue0']), dtype=torch.float64)
    for s in range(int(z['steps'])):
        x = torch.from_numpy(z[f'x{s}']).to(dev).requires_grad_(True)
        y = torch.from_numpy(z[f'y{s}']).to(dev).requires_grad_(True)
        xl, yl = torch.from_numpy(z[f'xl{s}']), torch.from_numpy(z[f'yl{s}'])
        # drive the two passes separately to read the bookkeeping of each (FFC.forward is the same two calls)
        px = F.normalize(x)
        loss2 = m.head(px, F.normalize(y).detach(), xl, yl, commit=False)
        rb = m.last_bookkeeping()
        py = F.normalize(y)
        loss1 = m.head(py, F.normalize(x).detach(), yl, xl, commit=True)
        cm = m.last_bookkeeping()
        loss = loss1 + loss2
        gx, gy = torch.autograd.grad(loss, [px, py])
        for got, pn in ((rb, 'rb'), (cm, 'cm')):
            for val, k in zip(got, ('rows', 'cols', 'labels', 'ones')):
                assert val == z[f'{pn}_{k}{s}'].tolist(), (s, pn, k)
        assert [list(kv) for kv in m.lru.state_dict()] == z[f'lru{s}'].tolist()
        qp = m.queue_position_dict
        assert [qp[i] for i in range(Q)] == z[f'qpos{s}'].tolist()
        # numerics: the golden (reference fp32) and the fp64 oracle
        x64 = torch.from_numpy(z[f'x{s}']).double().requires_grad_(True)
        y64 = torch.from_numpy(z[f'y{s}']).double().requires_grad_(True)
        l64 = o64.forward(x64, y64, xl.tolist(), yl.tolist())
        l64.backward()
        ref = float(z[f'loss{s}'])
        assert abs(float(loss) - float(l64)) <= tol_loss * abs(float(l64)), (s, float(loss), float(l64))
        assert abs(float(loss) - ref) <= max(tol_loss, 2e-5) * abs(ref), (s, float(loss), ref)
        assert _rel(gx.double().cpu(), x64.grad) <= tol_grad, (s, 'dx', _rel(gx.double().cpu(), x64.grad))
        assert _rel(gy.double().cpu(), y64.grad) <= tol_grad, (s, 'dy', _rel(gy.double().cpu(), y64.grad))
        assert _rel(gx.cpu(), torch.from_numpy(z[f'dx{s}'])) <= max(tol_grad, 3e-5)
    # queue after commit is a pure copy of the gallery embeddings
    assert np.abs(m.queue.cpu().numpy() - z['queue_final']).max() < 1e-6
    return m


@pytest.mark.parametrize('name', CASES)
def test_check_mode_matches_reference(golden_dir, name):
    """fp32 check mode: bookkeeping bit-exact, loss and gradients within 1e-5 relative."""
    z = np.load(os.path.join(golden_dir, f'ffc_{name}.npz'))
    _run_case(z, 'fp32', 1e-5, 1e-5)


@pytest.mark.parametrize('name', [c for c in CASES if c.endswith('d128')])
def test_bf16_tensor_core_matches_reference(golden_dir, name):
    """bf16 tcgen05 path: bookkeeping bit-exact, loss and gradients within 1e-2 relative."""
    z = np.load(os.path.join(golden_dir, f'ffc_{name}.npz'))
    _run_case(z, 'bf16', 1e-2, 1e-2)


def test_forward_module_and_grad_scaling():
    """FFC.forward (4-argument reference signature) + backward with a GradScaler-style upstream gradient."""
    import ffc_b200
    dev = torch.device('cuda')
    torch.manual_seed(3)
    D, Q, B = 64, 256, 32
    m = ffc_b200.FFC('identity', D, queue_size=Q, loss_type='Arc', margin=0.5, precision='fp32', max_batch=B).to(dev)
    o = HeadOracle(D, Q, 32.0, 'Arc', 0.5, queue=m.queue.cpu(), dtype=torch.float64)
    for step in range(3):
        x = torch.randn(B, D, device=dev, requires_grad=True)
        y = torch.randn(B, D, device=dev, requires_grad=True)
        xl = torch.randint(0, 300, (B,))
        yl = torch.cat([xl[:B // 2], torch.randint(0, 300, (B - B // 2,))])
        loss = m(x, y, xl, yl)
        (loss * 1024.0).backward()
        x64 = x.detach().cpu().double().requires_grad_(True)
        y64 = y.detach().cpu().double().requires_grad_(True)
        l64 = o.forward(F.normalize(x64), F.normalize(y64), xl.tolist(), yl.tolist())
        (l64 * 1024.0).backward()
        assert abs(float(loss) - float(l64)) <= 1e-5 * abs(float(l64))
        assert _rel(x.grad.cpu().double(), x64.grad) < 2e-5
        assert _rel(y.grad.cpu().double(), y64.grad) < 2e-5


def test_scatter_duplicates_last_wins():
    """q[r,c]=g with r=[0,0,1,0], c=[2,2,2,2], g=[1,2,3,4] -> q[0,2]=4, q[1,2]=3; restore undoes it."""
    from ffc_b200 import _capi
    lib = _capi.lib()
    dev = torch.device('cuda')
    Q, D = 5, 8
    q = torch.arange(2 * Q * D, dtype=torch.float32, device=dev).reshape(2, Q, D).contiguous()
    q0 = q.clone()
    qh = torch.zeros(2, Q, D, dtype=torch.bfloat16, device=dev)
    rows = torch.tensor([0, 0, 1, 0], dtype=torch.int32, device=dev)
    cols = torch.tensor([2, 2, 2, 2], dtype=torch.int32, device=dev)
    g = torch.tensor([1., 2., 3., 4.], device=dev).view(4, 1).expand(4, D).contiguous()
    undo = torch.zeros(4, D, device=dev)
    s = torch.cuda.current_stream().cuda_stream
    _capi.check(lib.ffc_queue_scatter(q.data_ptr(), qh.data_ptr(), rows.data_ptr(), cols.data_ptr(), g.data_ptr(), 4, Q, D, undo.data_ptr(), s))
    assert q[0, 2].tolist() == [4.0] * D and q[1, 2].tolist() == [3.0] * D
    assert qh[0, 2].float().tolist() == [4.0] * D and qh[1, 2].float().tolist() == [3.0] * D
    untouched = torch.ones(2, Q, dtype=torch.bool); untouched[0, 2] = False; untouched[1, 2] = False
    assert torch.equal(q.cpu()[untouched], q0.cpu()[untouched])
    _capi.check(lib.ffc_queue_restore(q.data_ptr(), qh.data_ptr(), rows.data_ptr(), cols.data_ptr(), undo.data_ptr(), 4, Q, D, s))
    assert torch.equal(q, q0)
