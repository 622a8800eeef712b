#!/usr/bin/env python
"""HBM roofline of the one-launch gallery EMA: 12 bytes per parameter (read g, read p, write g).
    python tools/ema_bench.py [n_params] [n_tensors]      default: 43.6 M parameters in 238 tensors (ir50-sized)"""
import json
import os
import sys

R = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [R, os.path.join(R, 'very-large-scale-face-recognition_b200')]
import torch
import torch.nn as nn
import ffc_b200

n_params = int(float(sys.argv[1])) if len(sys.argv) > 1 else 43_600_000
n_tensors = int(sys.argv[2]) if len(sys.argv) > 2 else 238
dev = torch.device('cuda')


class Bag(nn.Module):
    def __init__(self):
        super().__init__()
        g = torch.Generator().manual_seed(0)
        w = torch.rand(n_tensors, generator=g) ** 3 + 1e-3
        sizes = (w / w.sum() * n_params).long().clamp_min(1)
        self.ps = nn.ParameterList([nn.Parameter(torch.randn(int(s))) for s in sizes])

    def forward(self, x):
        return x


m = ffc_b200.FFC('x', 64, queue_size=128, probe_net=Bag(), gallery_net=Bag(), max_batch=16).to(dev)
n = sum(p.numel() for p in m.gallery_net.parameters())
for _ in range(3):
    m._momentum_update_gallery()
# time the kernel itself: direct C-ABI launches with the cached chunk table (the Python wrapper adds host time, not device time)
from ffc_b200 import _capi
lib = _capi.lib()
s = torch.cuda.current_stream().cuda_stream
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
ts = []
for _ in range(10):
    flush.zero_()                       # L2 flush between timed iterations
    e0.record()
    _capi.check(lib.ffc_ema_update(m._ema_table.data_ptr(), m._ema_chunks, 0.99, 0.01, s))
    e1.record()
    torch.cuda.synchronize()
    ts.append(e0.elapsed_time(e1))
ts.sort()
t = ts[len(ts) // 2]
peaks = json.load(open(os.path.join(R, 'MEASURED_PEAKS.json'))) if os.path.isfile(os.path.join(R, 'MEASURED_PEAKS.json')) else {'hbm_gbs': 6650.0}
gbs = 12.0 * n / (t * 1e-3) / 1e9
print(json.dumps(dict(kernel='ema_update_kernel', params=n, tensors=n_tensors, ms=t, algorithmic_bytes=12 * n, achieved_gbs=gbs, peak_gbs=peaks['hbm_gbs'],
                      frac=gbs / peaks['hbm_gbs'])))
