"""The BASELINE.json configurations themselves through the CUDA path (SURVEY.md 8, C1 / C2 / C4 head shapes):

  * C1 (batch 64, 10k identities, queue 4096, D = 128 / 512): the fixtures recorded from the unmodified reference
    (tests/golden/c1_*.npz) through ``ffc_b200.FFC`` in both precisions -- bookkeeping bit-exact, final queue bit-exact (SHA-256),
    loss / gradients within 1e-5 (fp32 check mode) and 1e-2 (bf16 tcgen05 path);
  * C2 shape (B = 512, queue 65 536, D = 512, 100k identities, k = 10): the bf16 path against the fp64 ORACLE (not against our own
    check mode), LRU full so that evictions occur;
  * C4 regime at a size the oracle holds (identities = 10 x queue, Zipf labels: most instance rows miss -> LRU eviction path, most
    probe labels unknown -> hard-negative rows dominate), bf16 against the fp64 oracle;
  * ArcFace's customary scale 64 (main.py:159 takes any --scale), both precisions.
"""
import glob
import hashlib
import os

import numpy as np
import pytest
import torch
import torch.nn.functional as F

from oracle.head_ref import HeadOracle

pytestmark = pytest.mark.gpu

C1_CASES = sorted(os.path.basename(p)[3:-4] for p in glob.glob(os.path.join(os.path.dirname(__file__), 'golden', 'c1_*.npz')))


def _rel(a, b):
    return float((a - b).norm() / (b.norm() + 1e-30))


def seeded_queue(Q, D, seed):
    return F.normalize(torch.rand(2, Q, D, generator=torch.Generator().manual_seed(seed)), dim=2)


@pytest.mark.parametrize('precision,tol', [('fp32', 1e-5), ('bf16', 1e-2)])
@pytest.mark.parametrize('name', C1_CASES)
def test_c1_fixtures_through_the_cuda_head(golden_dir, name, precision, tol):
    """ffc.py:264-267 at main.py:151-163's shapes: what the reference computed on CPU, step by step."""
    import ffc_b200
    z = np.load(os.path.join(golden_dir, f'c1_{name}.npz'))
    dev = torch.device('cuda')
    D, Q, B = int(z['D']), int(z['Q']), int(z['B'])
    assert (Q, B, int(z['n_ids'])) == (4096, 64, 10000)
    m = ffc_b200.FFC('identity', D, queue_size=Q, scale=float(z['scale']), loss_type=str(z['loss_type']), margin=float(z['margin']),
                     precision=precision, max_batch=B)
    m.queue.copy_(seeded_queue(Q, D, int(z['queue_seed'])))
    m = m.to(dev)
    m.lru.restore([(i, i) for i in range(int(z['warm']))])
    for s in range(int(z['steps'])):
        # the reference's stand-in backbone re-normalises its (unit-norm) input on the CPU; feed the head those very rows so that the
        # enqueued rows -- pure copies -- can be pinned by the reference's hash
        px = F.normalize(torch.from_numpy(z[f'x{s}'])).to(dev).requires_grad_(True)
        py = F.normalize(torch.from_numpy(z[f'y{s}'])).to(dev).requires_grad_(True)
        xl, yl = torch.from_numpy(z[f'xl{s}']), torch.from_numpy(z[f'yl{s}'])
        loss2 = m.head(px, py.detach(), xl, yl, commit=False)
        rb = m.last_bookkeeping()
        loss1 = m.head(py, px.detach(), yl, xl, commit=True)
        cm = m.last_bookkeeping()
        gx, gy = torch.autograd.grad(loss1 + loss2, [px, py])
        for got, pn in ((rb, 'rb'), (cm, 'cm')):
            for val, k in zip(got, ('rows', 'cols', 'labels', 'ones')):
                assert val == z[f'{pn}_{k}{s}'].tolist(), (s, pn, k)
        assert [list(kv) for kv in m.lru.state_dict()] == z[f'lru{s}'].tolist(), (s, 'lru')
        qp = m.queue_position_dict                        # (a property: one device read per access)
        assert [qp[i] for i in range(Q)] == z[f'qpos{s}'].tolist(), (s, 'qpos')
        ref = float(z[f'loss{s}'])
        assert abs(float(loss1 + loss2) - ref) <= max(tol, 2e-5) * abs(ref), (s, float(loss1 + loss2), ref)
        assert _rel(gx.cpu(), torch.from_numpy(z[f'dx{s}'])) <= max(tol, 3e-5), (s, 'dx')
        assert _rel(gy.cpu(), torch.from_numpy(z[f'dy{s}'])) <= max(tol, 3e-5), (s, 'dy')
    sha = hashlib.sha256(m.queue.detach().cpu().contiguous().numpy().tobytes()).hexdigest()
    assert sha == str(z['queue_final_sha256'])                      # the [2, 4096, D] queue after 5 steps, bit for bit


def _center(i, D, seed):
    return torch.randn(D, generator=torch.Generator().manual_seed(seed * 1000003 + int(i)))


def _clustered(ids, D, seed, noise, gen):
    c = torch.stack([_center(i, D, seed) for i in ids.tolist()])
    return F.normalize(F.normalize(c) + noise * torch.randn(len(ids), D, generator=gen) / D ** 0.5)


def _zipf(n, N, a, gen):
    """n draws from a truncated Zipf(a) over [0, N) (inverse-CDF on a seeded uniform), popular ids first"""
    u = torch.rand(n, generator=gen, dtype=torch.float64)
    x = ((N ** (1 - a) - 1) * u + 1) ** (1 / (1 - a))
    return (x.floor().long() - 1).clamp_(0, N - 1)


def _run_vs_oracle(D, Q, B, N, loss_type, margin, scale, steps, label_fn, seed, tol=1e-2, check_full_lru=True, prefill_offset=0):
    import ffc_b200
    dev = torch.device('cuda')
    q0 = seeded_queue(Q, D, seed)
    h = ffc_b200.FFCHead(D, Q, scale, loss_type, margin, precision='bf16', max_batch=B, device=dev)
    h.queue.copy_(q0.to(dev))
    h._ensure()
    h.sync_mirror()
    o = HeadOracle(D, Q, scale, loss_type, margin, queue=q0, dtype=torch.float64)
    # LRU full (ids prefill_offset .. prefill_offset + Q - 1 resident): every miss evicts
    h.lru.restore_arrays(torch.arange(prefill_offset, prefill_offset + Q, dtype=torch.int64), torch.arange(Q, dtype=torch.int32))
    o.lru.restore([(prefill_offset + i, i) for i in range(Q)])
    gen = torch.Generator().manual_seed(seed + 1)
    seen = dict(out=0, pos=0, evict=0, ones=0)
    for s in range(steps):
        xl, yl = label_fn(gen)
        x = _clustered(xl, D, seed, 0.8, gen)
        y = _clustered(yl, D, seed, 0.8, gen)
        xd, yd = x.to(dev), y.to(dev)
        loss, dx, dy = h.forward_pair(xd, yd, yd, xd, xl, yl)
        x64, y64 = x.double().requires_grad_(True), y.double().requires_grad_(True)
        ref = o.forward(x64, y64, xl.tolist(), yl.tolist())
        ref.backward()
        # bookkeeping of both passes, bit for bit (sets: 0 = rollback pass, 1 = commit pass)
        torch.cuda.synchronize()
        for st, tr in zip(h._sets, o.trace[-2:]):
            n1 = int(st['n_ones'].item())
            assert st['rows'][:B].tolist() == tr['rows'] and st['cols'][:B].tolist() == tr['cols'], s
            assert st['label'][:B].tolist() == tr['labels'], s
            assert sorted(st['ones_list'][:n1].tolist()) == tr['ones'], s
        for tr in o.trace[-2:]:
            seen['out'] += sum(l < 0 for l in tr['labels'])
            seen['pos'] += sum(l >= 0 for l in tr['labels'])
            seen['ones'] += len(tr['ones'])
        assert abs(float(loss) - float(ref)) <= tol * abs(float(ref)), (s, float(loss), float(ref))
        assert _rel(dx.double().cpu(), x64.grad) <= tol, (s, 'dx', _rel(dx.double().cpu(), x64.grad))
        assert _rel(dy.double().cpu(), y64.grad) <= tol, (s, 'dy', _rel(dy.double().cpu(), y64.grad))
    if check_full_lru:
        assert h.lru.state_dict() == o.lru.state_dict()
    qp = h.queue_position_dict
    assert [qp[i] for i in range(Q)] == o.qpos
    assert torch.equal(h.queue.cpu(), o.queue.float())                     # enqueue is a pure copy
    return seen


def test_c2_shape_bf16_against_the_oracle():
    """BASELINE.json configs[1]'s head: batch 512, 100k identities, queue 65 536, D = 512 (k = hard_neg = 10, ffc.py:48)."""
    D, Q, B, N = 512, 65536, 512, 100000
    perm = torch.randperm(N, generator=torch.Generator().manual_seed(5))
    it = [0]

    def labels(gen):
        h = B // 2
        ids = perm[it[0] * h:(it[0] + 1) * h]
        it[0] += 1
        return (torch.cat([ids, torch.randint(0, N, (B - h,), generator=gen)]), torch.cat([ids, torch.randint(0, N, (B - h,), generator=gen)]))
    seen = _run_vs_oracle(D, Q, B, N, 'Arc', 0.5, 32.0, 2, labels, seed=11)
    assert seen['out'] > 50 and seen['pos'] > 300 and seen['ones'] > 100        # outliers, known targets and `ones` slots all occur


@pytest.mark.parametrize('draw', ['uniform', 'zipf'])
def test_c4_regime_eviction_and_outlier_heavy_bf16_against_the_oracle(draw):
    """C4's regime (identities = 10 x queue, LRU-managed, the LRU full of OTHER identities): every miss evicts (lru.py:74-89, also
    inside a batch) and the probe labels of the instance half are mostly unknown (ffc.py:86-92: hard-negative rows).  `uniform` is
    bench.py's c4 draw (nearly every instance row an outlier), `zipf` a skewed one (popular identities become resident).  k = 10."""
    D, Q, B = 512, 50176, 512
    N = 10 * Q

    def labels(gen):
        h = B // 2
        ids = torch.randperm(N, generator=gen)[:h]          # id half: distinct identities, mostly never seen -> misses
        if draw == 'uniform':
            a, b = torch.randint(0, N, (B - h,), generator=gen), torch.randint(0, N, (B - h,), generator=gen)
        else:
            a, b = _zipf(B - h, N, 1.1, gen), _zipf(B - h, N, 1.1, gen)
        return torch.cat([ids, a]), torch.cat([ids, b])
    seen = _run_vs_oracle(D, Q, B, N, 'Arc', 0.5, 32.0, 3, labels, seed=21, prefill_offset=N)
    rows = seen['out'] + seen['pos']
    assert seen['out'] >= (0.4 if draw == 'uniform' else 0.12) * rows, seen          # outlier-heavy
    assert seen['pos'] > 0.3 * rows


@pytest.mark.parametrize('loss_type,margin', [('Arc', 0.5), ('AM', 0.4)])
def test_scale_64(loss_type, margin):
    """main.py:159 takes any --scale; ArcFace's customary s = 64: the fixed softmax reference point is centred on the logit range,
    so nothing under- or overflows (csrc/head.cu fixed_max_of)."""
    import ffc_b200
    dev = torch.device('cuda')
    D, Q, B, n_ids = 128, 3000, 96, 4000
    gen = torch.Generator().manual_seed(3)
    for precision, tol in (('fp32', 1e-5), ('bf16', 1e-2)):
        m = ffc_b200.FFC('identity', D, queue_size=Q, scale=64.0, loss_type=loss_type, margin=margin, precision=precision, max_batch=B).to(dev)
        o = HeadOracle(D, Q, 64.0, loss_type, margin, queue=m.queue.cpu(), dtype=torch.float64)
        for step in range(3):
            ids = torch.randperm(n_ids, generator=gen)[:B // 2]
            xl = torch.cat([ids, torch.randint(0, n_ids, (B - B // 2,), generator=gen)])
            yl = torch.cat([ids, torch.randint(0, n_ids, (B - B // 2,), generator=gen)])
            x = _clustered(xl, D, 9, 0.7, gen)
            y = _clustered(yl, D, 9, 0.7, gen)
            xd, yd = x.to(dev).requires_grad_(True), y.to(dev).requires_grad_(True)
            loss = m(xd, yd, xl, yl)
            loss.backward()
            x64, y64 = x.double().requires_grad_(True), y.double().requires_grad_(True)
            ref = o.forward(F.normalize(x64), F.normalize(y64), xl.tolist(), yl.tolist())
            ref.backward()
            assert m.lru.state_dict() == o.lru.state_dict()
            assert abs(float(loss) - float(ref)) <= tol * abs(float(ref)), (precision, step, float(loss), float(ref))
            assert _rel(xd.grad.double().cpu(), x64.grad) <= 2 * tol and _rel(yd.grad.double().cpu(), y64.grad) <= 2 * tol, (precision, step)


def test_scale_64_d512_bf16():
    """the same at D = 512 with a queue large enough for k = 10 and several column chunks"""
    D, Q, B, N = 512, 50176, 256, 60000

    def labels(gen):
        h = B // 2
        ids = torch.randperm(N, generator=gen)[:h]
        return (torch.cat([ids, torch.randint(0, N, (B - h,), generator=gen)]), torch.cat([ids, torch.randint(0, N, (B - h,), generator=gen)]))
    _run_vs_oracle(D, Q, B, N, 'Arc', 0.5, 64.0, 2, labels, seed=31)
