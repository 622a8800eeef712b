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
    assert r.returncode == 0 and 'DIST_GPU_OK' in r.stdout, r.stdout[-3000:] + r.stderr[-3000:]
