"""The persistent grid of the CTA-pair sweep (csrc/head_sm100.cu): with more than 2 x 74 work items a launch is 74 CTA pairs that walk
the item list -- barrier phases, TMEM and the pipeline rings run on across items, the O write-out is staged through the P~ buffer of the
item's last tile.  The parity suites normally run one pair per item (their shapes have few items); here they are re-run in a subprocess
with the grid forced down to one or two pairs (FFC_SWEEP_PAIRS, read once per process), so that every launch walks several items per
pair: main / side items, items without columns, ragged row and column tiles, several column chunks.  FFC_SWEEP_ROWMAP=1 additionally
forces the "hard-negative-only rows last" row order (normally used only by launches several waves deep), so that the fixtures with unknown
probe labels run row tiles without exponentials / GEMM-2 and seed their top-k lists from finished items.  And one many-row-tile shape
against the fp64 check mode."""
import os
import subprocess
import sys

import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.parametrize('env', [dict(FFC_SWEEP_PAIRS='1'), dict(FFC_SWEEP_PAIRS='2', FFC_SWEEP_CHUNKS='3', FFC_SWEEP_ROWMAP='1'), dict(FFC_SWEEP_ROWMAP='1')],
                         ids=['one_pair', 'two_pairs_three_chunks_rowmap', 'rowmap'])
def test_parity_suites_on_a_forced_persistent_grid(env):
    # (tests/test_gpu_fast_paths.py is left out: it compares sharded against unsharded runs to 2e-5, which presumes both use the same
    # column-chunk split; the overrides change the split, i.e. the fp32 summation order)
    files = ['tests/test_gpu_head.py', 'tests/test_gpu_head_shapes.py', 'tests/test_gpu_baseline_configs.py', 'tests/test_gpu_dqueue.py']
    r = subprocess.run([sys.executable, '-m', 'pytest', '-q', '-x', '-m', 'gpu', '-k', 'not fp32 and not check_mode'] + files, cwd=ROOT, env=dict(os.environ, **env),
                       capture_output=True, text=True, timeout=1500)
    assert r.returncode == 0, (r.stdout + r.stderr)[-4000:]


def _rel(a, b):
    return float((a - b).norm() / (b.norm() + 1e-30))


def test_many_row_tiles_persistent_by_shape():
    """2304 probe rows (18 row tiles) x 65 536 columns, D = 512: more items than 2 x 74, so the launch is persistent without any
    override.  bf16 against the fp64 check mode on the same state, both passes, hits / fresh inserts / outliers mixed."""
    import ffc_b200
    dev = torch.device('cuda')
    Q, D, B = 65536, 512, 2304
    hb = ffc_b200.FFCHead(D, Q, 32.0, 'Arc', 0.5, precision='bf16', max_batch=B, device=dev)
    hc = ffc_b200.FFCHead(D, Q, 32.0, 'Arc', 0.5, precision='fp32', max_batch=B, device=dev)
    hc.queue.copy_(hb.queue)
    for h in (hb, hc):
        h._ensure()
        h.sync_mirror()
        n0 = 3 * Q // 4
        h._lru.restore_arrays(torch.arange(n0, dtype=torch.int64), torch.arange(n0, dtype=torch.int32))
    gen = torch.Generator().manual_seed(3)
    for step in range(2):
        xl = torch.randint(0, Q + Q // 8, (B,), generator=gen)
        yl = torch.cat([xl[:B // 2], torch.randint(0, Q + Q // 8, (B - B // 2,), generator=gen)])
        x = F.normalize(torch.randn(B, D, generator=gen)).to(dev)
        y = F.normalize(torch.randn(B, D, generator=gen)).to(dev)
        out = []
        for h in (hb, hc):
            l2, d2 = h._pass(x, y, xl, yl, False)
            l1, d1 = h._pass(y, x, yl, xl, True)
            out.append((l1 + l2, d1, d2, h.label[:B].clone()))
        (lb, d1b, d2b, lab_b), (lc, d1c, d2c, lab_c) = out
        assert torch.equal(lab_b, lab_c)
        assert int((lab_b < 0).sum()) > 0 and int((lab_b >= 0).sum()) > 0
        assert abs(float(lb) - float(lc)) <= 1e-2 * abs(float(lc)), (float(lb), float(lc))
        assert _rel(d1b, d1c) <= 1e-2 and _rel(d2b, d2c) <= 1e-2, (_rel(d1b, d1c), _rel(d2b, d2c))
