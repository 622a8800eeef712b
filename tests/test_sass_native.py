"""SURVEY.md section 7 test matrix (v): the built library's SASS for sm_100a carries the Blackwell-native instructions the design claims --
tcgen05.mma (`UTCHMMA`), tcgen05.ld / st (`LDTM` / `STTM`), TMA tensor loads (`UTMALDG`), mbarrier waits (`SYNCS`) -- in the sweep kernel,
and no `HMMA` (mma.sync / wmma) fallback.  Runs on the build host: cuobjdump needs no GPU."""
import os
import re
import shutil
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CUOBJDUMP = shutil.which('cuobjdump') or '/usr/local/cuda/bin/cuobjdump'


@pytest.fixture(scope='module')
def sass():
    if not os.path.isfile(CUOBJDUMP):
        pytest.skip('cuobjdump not available')
    from ffc_b200 import _capi
    import importlib.util
    spec = importlib.util.spec_from_file_location('ffc_build', os.path.join(ROOT, 'very-large-scale-face-recognition_b200', 'build.py'))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    mod.build()
    r = subprocess.run([CUOBJDUMP, '-sass', _capi.LIB_PATH], capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stderr[-2000:]
    return r.stdout


def _functions(sass):
    out, name = {}, None
    for line in sass.splitlines():
        m = re.search(r'Function : (\S+)', line)
        if m:
            name = m.group(1)
            out[name] = []
        elif name is not None:
            out[name].append(line)
    return {k: '\n'.join(v) for k, v in out.items()}


def test_every_kernel_is_sm_100a(sass):
    """each fatbin section that holds code is sm_100a (the only other section is nvcc's empty device-link stub of the `-shared` step)"""
    sections = re.split(r'Fatbin elf code:', sass)[1:]
    with_code = [s for s in sections if 'Function :' in s]
    assert len(with_code) >= 6
    for s in with_code:
        assert re.search(r'arch = (sm_\w+)', s).group(1) == 'sm_100a'
    for s in sections:
        if 'Function :' not in s:
            assert 'EXIT' not in s                               # no instructions at all


def test_sweep_kernel_uses_tcgen05_tmem_tma(sass):
    fns = _functions(sass)
    sweeps = {k: v for k, v in fns.items() if 'ffc_head_sweep_sm100_kernel' in k}
    assert len(sweeps) >= 3                                     # templated on D: 128 / 256 / 512
    for name, body in sweeps.items():
        for mnemonic in ('UTCHMMA', 'LDTM', 'UTMALDG', 'SYNCS'):
            assert mnemonic in body, (name, mnemonic)
        assert 'tmem[' in body, name                            # TMEM operands (accumulators; probe tile as the A operand of GEMM-1)
        assert not re.search(r'\bHMMA\b', body), name           # no mma.sync / wmma path
    assert any('STTM' in b for b in sweeps.values())            # tcgen05.st: probe tile into TMEM


def test_one_cta_sweep_keeps_both_operands_of_gemm2_out_of_shared_memory(sass):
    """the D <= 256 kernel (csrc/head_sm100_1cta.cu): tcgen05.mma with TMEM A operands for BOTH GEMMs, p~ written back with
    tcgen05.st (STTM.x16) over the S columns -- no st.async / DSMEM hand-off (STAS) and no cluster barrier traffic"""
    fns = _functions(sass)
    ones = {k: v for k, v in fns.items() if 'ffc_head_sweep1_sm100_kernel' in k}
    assert len(ones) >= 6                                       # {SV, not SV} x D 64 / 128 / 256
    for name, body in ones.items():
        for mnemonic in ('UTCHMMA', 'LDTM', 'STTM', 'UTMALDG', 'SYNCS'):
            assert mnemonic in body, (name, mnemonic)
        assert 'STTM.x16' in body or 'STTM.16' in body or re.search(r'STTM\S*x16', body), name
        assert 'STAS' not in body and not re.search(r'\bHMMA\b', body), name
        assert len(re.findall(r'UTCHMMA', body)) >= 8, name


def test_every_kernel_family_is_in_the_library(sass):
    fns = ' '.join(_functions(sass))
    for family in ('lru_', 'queue_scatter_kernel', 'queue_restore_kernel', 'route_keys', 'head_prep_fused_kernel', 'head_finalize_fused_kernel',
                   'head_sweep_simt_kernel', 'ema_update_kernel', 'tail_rows_fwd_kernel', 'tail_col_stats_kernel', 'tail_rows_bwd_bn_kernel',
                   'tail_cols_dx_kernel', 'head_thr_from_tgt_kernel', 'sum_slabs_barrier_kernel', 'overlay_clear_kernel', 'dq_special_rows_kernel',
                   'dq_hard_neg_kernel', 'ffc_head_sweep1_sm100_kernel'):
        assert family in fns, family
