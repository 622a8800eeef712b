"""GPU parity: device LRU vs the oracle (oracle/lru_ref.py) and the reference traces.  Bit-exact."""
import json
import os
import random

import pytest
import torch

from oracle.lru_ref import LRU as RefLRU

pytestmark = pytest.mark.gpu


def _dev_lru(cap, **kw):
    import ffc_b200
    return ffc_b200.LRU(cap, **kw)


def test_reference_traces_scalar_api(golden_dir):
    with open(os.path.join(golden_dir, 'lru_traces.json')) as f:
        cases = json.load(f)
    n = 0
    for case in cases:
        lru = _dev_lru(case['capacity'])
        if 'restore' in case:
            lru.restore([tuple(kv) for kv in case['restore']])
            assert [lru.get(k) for k in case['gets']] == case['slots']
            assert [list(kv) for kv in lru.state_dict()] == case['final']
            assert sorted(lru.keys()) == sorted(k for k, _ in case['final'])
            continue
        ops = case['ops'] if case['capacity'] <= 16 else case['ops'][:400]
        for op in ops:
            kind, arg = op[0], op[1]
            if kind == 'get':
                assert lru.get(arg) == op[2], (case['capacity'], n)
            elif kind == 'try_get':
                assert lru.try_get(arg) == op[2], (case['capacity'], n)
            elif kind == 'rollback_steps':
                lru.rollback_steps(arg)
            elif kind == 'view':
                assert lru.view(arg) == op[2]
            elif kind == 'contains':
                assert (arg in lru) == op[2]
            else:
                assert [list(kv) for kv in lru.state_dict()] == op[2]
                assert lru.cur_idx == op[3]
            n += 1
    assert n > 1000


def _ref_batch(ref, qpos, keys, journal):
    """ffc.py:166-177 over the oracle LRU: returns rows, cols, hits, ones(set)"""
    rows, cols, hits, ones = [], [], [], set()
    for k in keys:
        known = k in ref
        s = ref.try_get(k) if journal else ref.get(k)
        if known:
            rows.append(qpos[s]); ones.add(s); qpos[s] ^= 1
        else:
            rows.append(0); qpos[s] = 1
        cols.append(s); hits.append(1 if known else 0)
    return rows, cols, hits, ones


@pytest.mark.parametrize('cap,universe,batch', [(1, 5, 7), (2, 6, 16), (7, 12, 16), (64, 90, 48), (64, 4000, 64),
                                                 (1000, 1500, 256), (1000, 100000, 512), (5000, 5200, 1024), (300, 100000, 1024)])
def test_batched_assign_matches_oracle(cap, universe, batch):
    dev = torch.device('cuda')
    rng = random.Random(cap * 7919 + batch)
    lru = _dev_lru(cap)
    ref = RefLRU(cap)
    qpos = torch.zeros(cap, dtype=torch.uint8, device=dev)
    ref_qpos = [0] * cap
    cmask = torch.zeros(cap // 32 + 2, dtype=torch.int32, device=dev)
    for step in range(14):
        n = batch if step % 3 else rng.randrange(1, batch + 1)
        # mix of fresh ids, repeats inside the batch and old ids
        keys = [rng.randrange(universe) for _ in range(n)]
        if step % 4 == 1:
            keys = [keys[rng.randrange(max(1, n // 3))] for _ in range(n)]
        journal = step % 2 == 1
        kt = torch.tensor(keys, dtype=torch.int64, device=dev)
        rows = torch.empty(n, dtype=torch.int32, device=dev)
        cols = torch.empty(n, dtype=torch.int32, device=dev)
        hit = torch.empty(n, dtype=torch.uint8, device=dev)
        ones_list = torch.empty(n, dtype=torch.int32, device=dev)
        n_ones = torch.zeros(1, dtype=torch.int32, device=dev)
        before = ref.state_dict()
        q_before = list(ref_qpos)
        lru.assign(kt, journal=journal, qpos=qpos, rows=rows, cols=cols, hit=hit, ones_list=ones_list, n_ones=n_ones, cmask=cmask)
        r_rows, r_cols, r_hits, r_ones = _ref_batch(ref, ref_qpos, keys, journal)
        assert cols.tolist() == r_cols, (step, journal)
        assert rows.tolist() == r_rows
        assert hit.tolist() == r_hits
        no = int(n_ones.item())
        assert sorted(ones_list[:no].tolist()) == sorted(r_ones)
        bits = set()
        for w, v in enumerate(cmask.tolist()):
            v &= 0xffffffff
            while v:
                b = (v & -v).bit_length() - 1
                bits.add(w * 32 + b)
                v &= v - 1
        assert bits == r_ones
        cmask.zero_()
        # probe view after the batch
        probe = [rng.randrange(universe) for _ in range(37)] + keys[:5]
        pv = lru.view_batch(torch.tensor(probe, dtype=torch.int64, device=dev)).tolist()
        assert pv == [ref.view(k) for k in probe]
        assert lru.state_dict() == ref.state_dict()
        assert qpos.cpu().tolist() == ref_qpos
        if journal:
            if step % 4 == 3 and n > 2:      # partial rollback first (lru.py:252-255 semantics)
                part = rng.randrange(1, n)
                lru.undo(part, qpos)
                ref.rollback_steps(part)
                assert lru.state_dict() == ref.state_dict()
                lru.undo(n - part, qpos)
                ref.rollback_steps(n - part)
            else:
                lru.undo(n, qpos)
                ref.rollback_steps(n)
            ref_qpos[:] = q_before
            assert lru.state_dict() == before
            assert qpos.cpu().tolist() == q_before
            assert lru.cur_idx == ref.cur_idx
            assert lru.journal_len == 0


def test_long_run_compaction_and_rebuild():
    """Enough batches to force ring compaction and hash-table rebuilds (capacity 512 -> ring 16384)."""
    dev = torch.device('cuda')
    rng = random.Random(5)
    cap = 512
    lru, ref = _dev_lru(cap), RefLRU(cap)
    for step in range(80):
        universe = 600 if step % 2 else 100000
        keys = [rng.randrange(universe) for _ in range(1024)]
        cols = lru.assign(torch.tensor(keys, dtype=torch.int64, device=dev))
        assert cols.tolist() == [ref.get(k) for k in keys], step
    assert lru.state_dict() == ref.state_dict()


def test_export_import_round_trip():
    dev = torch.device('cuda')
    rng = random.Random(9)
    a, ref = _dev_lru(100), RefLRU(100)
    keys = [rng.randrange(400) for _ in range(700)]
    a.assign(torch.tensor(keys, dtype=torch.int64, device=dev))
    for k in keys:
        ref.get(k)
    sd = a.state_dict()
    assert sd == ref.state_dict()
    b = _dev_lru(100)
    b.restore(sd)
    ref2 = RefLRU(100)
    ref2.restore(sd)
    more = [rng.randrange(500) for _ in range(300)]
    assert b.assign(torch.tensor(more, dtype=torch.int64, device=dev)).tolist() == [ref2.get(k) for k in more]
    assert b.state_dict() == ref2.state_dict()
    b.clear()
    assert b.state_dict() == [] and b.cur_idx == 0
    assert b.get(77) == 0
