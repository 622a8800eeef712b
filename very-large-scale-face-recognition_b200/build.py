"""Build libffc_b200.so in-tree with nvcc for sm_100a (no JIT cache: the .so travels with the repo snapshot)."""
from __future__ import annotations

import concurrent.futures
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, 'csrc')
OUT_DIR = os.path.join(HERE, 'ffc_b200')
LIB = os.path.join(OUT_DIR, 'libffc_b200.so')
OBJ_DIR = os.path.join(HERE, 'build')
NVCC = os.environ.get('NVCC', '/usr/local/cuda/bin/nvcc')
FLAGS = ['-gencode', 'arch=compute_100a,code=sm_100a', '-O3', '-lineinfo', '-std=c++17', '-Xcompiler', '-fPIC',
         '--expt-relaxed-constexpr', '-I', os.path.join(HERE, '..', 'include')]
if os.environ.get('FFC_SM100_DEBUG_BUILD') == '1':      # bottleneck-isolation switches of the sweep (tools/sweep_modes.py); never shipped
    FLAGS.append('-DFFC_SM100_DEBUG_BUILD=1')


def _sources():
    return sorted(f for f in os.listdir(CSRC) if f.endswith('.cu'))


def _newest_input():
    paths = [os.path.join(CSRC, f) for f in os.listdir(CSRC)] + [os.path.join(HERE, '..', 'include', 'ffc_b200.h')]
    return max(os.path.getmtime(p) for p in paths)


def build(force: bool = False, verbose: bool = False) -> str:
    if not force and os.path.isfile(LIB) and os.path.getmtime(LIB) >= _newest_input():
        return LIB
    os.makedirs(OBJ_DIR, exist_ok=True)
    objs = []

    def compile_one(src):
        obj = os.path.join(OBJ_DIR, src[:-3] + '.o')
        cmd = [NVCC] + FLAGS + (['-Xptxas', '-v'] if verbose else []) + ['-c', os.path.join(CSRC, src), '-o', obj]
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError(f'nvcc failed on {src}:\n{r.stdout}\n{r.stderr}')
        if verbose:
            sys.stderr.write(r.stderr)
        return obj

    with concurrent.futures.ThreadPoolExecutor(max_workers=8) as ex:
        objs = list(ex.map(compile_one, _sources()))
    cmd = [NVCC, '-shared', '-o', LIB] + objs + ['-cudart', 'static']
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError(f'link failed:\n{r.stdout}\n{r.stderr}')
    return LIB


def build_debug_lib() -> str:
    """The bottleneck-isolation build of the sweep (-DFFC_SM100_DEBUG_BUILD=1: per-role cycle counters, switch-off modes) as a SEPARATE
    library, build_dbg/libffc_b200_dbg.so; select it with FFC_B200_LIB=<path> (tools/sweep_modes.py).  Never the shipped library."""
    out_dir = os.path.join(HERE, 'build_dbg')
    os.makedirs(out_dir, exist_ok=True)
    lib = os.path.join(out_dir, 'libffc_b200_dbg.so')
    objs = []
    for src in _sources():
        obj = os.path.join(out_dir, src[:-3] + '.o')
        r = subprocess.run([NVCC] + FLAGS + ['-DFFC_SM100_DEBUG_BUILD=1', '-c', os.path.join(CSRC, src), '-o', obj], capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError(f'nvcc failed on {src}:\n{r.stdout}\n{r.stderr}')
        objs.append(obj)
    r = subprocess.run([NVCC, '-shared', '-o', lib] + objs + ['-cudart', 'static'], capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError(f'link failed:\n{r.stdout}\n{r.stderr}')
    return lib


if __name__ == '__main__':
    if '--debug-lib' in sys.argv:
        print(build_debug_lib())
    else:
        print(build(force='--force' in sys.argv, verbose='-v' in sys.argv))
