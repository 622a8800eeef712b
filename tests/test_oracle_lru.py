"""Pin oracle/lru_ref.py against traces recorded from the reference lru.py (tests/golden/lru_traces.json)."""
import json
import os

import pytest

from oracle.lru_ref import LRU


def _cases(golden_dir):
    with open(os.path.join(golden_dir, 'lru_traces.json')) as f:
        return json.load(f)


def test_kat_from_survey():
    # known-answer vector probed on the reference.
    l = LRU(4)
    assert [l.get(k) for k in (10, 11, 12, 13, 10, 14, 15, 11)] == [0, 1, 2, 3, 0, 1, 2, 3]
    assert l.state_dict() == [(11, 3), (15, 2), (14, 1), (10, 0)]
    assert l.view(12) == -1 and l.view(10) == 0
    assert [l.try_get(k) for k in (99, 10, 98)] == [0, 1, 2]
    assert l.state_dict() == [(98, 2), (10, 1), (99, 0), (11, 3)]
    l.rollback_steps(3)
    assert l.state_dict() == [(11, 3), (15, 2), (14, 1), (10, 0)] and l.cur_idx == 4
    l2 = LRU(2)
    assert [l2.get(k) for k in (5, 6, 7, 5, 6)] == [0, 1, 0, 1, 0]


def test_traces_match_reference(golden_dir):
    n = 0
    for case in _cases(golden_dir):
        if 'ops' not in case:
            continue
        lru = LRU(case['capacity'])
        for op in case['ops']:
            kind, arg = op[0], op[1]
            if kind == 'get':
                assert lru.get(arg) == op[2]
            elif kind == 'try_get':
                assert lru.try_get(arg) == op[2]
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
    assert n > 3000


def test_restore_round_trip(golden_dir):
    case = [c for c in _cases(golden_dir) if 'restore' in c][0]
    lru = LRU(case['capacity'])
    lru.restore([tuple(kv) for kv in case['restore']])
    assert [lru.get(k) for k in case['gets']] == case['slots']
    assert [list(kv) for kv in lru.state_dict()] == case['final']
    assert sorted(lru.keys()) == sorted(k for k, _ in case['final'])
    assert [list(kv) for kv in lru] == case['final']
