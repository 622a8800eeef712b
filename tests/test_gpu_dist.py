"""Multi-GPU parity (NCCL): R-way sharded head == dense oracle with sharded bookkeeping.  Needs >= 2 GPUs."""
import os
import socket
import subprocess
import sys

import pytest
import torch

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize('world', [2, 8])
def test_sharded_head_nccl(world):
    if torch.cuda.device_count() < world:
        pytest.skip(f'needs {world} GPUs')
    with socket.socket() as s:
        s.bind(('127.0.0.1', 0))
        port = s.getsockname()[1]
    worker = os.path.join(os.path.dirname(os.path.abspath(__file__)), 'dist_gpu_worker.py')
    r = subprocess.run([sys.executable, '-m', 'torch.distributed.run', '--nnodes=1', f'--nproc-per-node={world}', '--master-addr', '127.0.0.1',
                        '--master-port', str(port), worker], capture_output=True, text=True, timeout=600)
    if r.returncode != 0:       # keep the whole worker log (the first failing rank's message scrolls out of a tail)
        out_dir = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), 'gpurun_out')
        if os.path.isdir(out_dir):
            with open(os.path.join(out_dir, f'dist_worker_world{world}.log'), 'w') as f:
                f.write(r.stdout + '\n---- stderr ----\n' + r.stderr)
    assert r.returncode == 0 and 'DIST_GPU_OK' in r.stdout, '\n'.join(l for l in (r.stdout + r.stderr).splitlines() if 'rank' in l and 'File' not in l)[-4000:]
