#!/usr/bin/env python
"""Bottleneck isolation of the tcgen05 sweep: run bench.py once per FFC_SM100_DEBUG bitmask and print the main-sweep time.
    python tools/sweep_modes.py 0 32 16 ...   (see Sm100Params::debug in csrc/head_sm100.cu)"""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
modes = sys.argv[1:] or ['0']
extra = os.environ.get('SWEEP_BENCH_ARGS', '--steps 12 --warmup 3 --no-cpu').split()
for m in modes:
    env = dict(os.environ, FFC_SM100_DEBUG=m)
    r = subprocess.run([sys.executable, os.path.join(ROOT, 'bench.py')] + extra, env=env, capture_output=True, text=True)
    if os.environ.get('FFC_SM100_DUMP'):
        print(r.stderr[-3000:], flush=True)
    try:
        d = json.loads(r.stdout.strip().splitlines()[-1])
        print(f"mode {m:>3}: sweep {d['roofline']['avg_ms']:.3f} ms  {d['roofline']['achieved']:.0f} TF  step {d['ms_per_step']:.3f} ms  "
              f"loss {d['loss']:.4f}  clocks {d['clocks']['sm_mhz']} {d['clocks']['reasons']}", flush=True)
    except Exception as e:
        print(f'mode {m}: FAILED rc={r.returncode} {e}\n{r.stdout[-500:]}\n{r.stderr[-1500:]}', flush=True)
