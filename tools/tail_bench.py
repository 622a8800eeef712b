#!/usr/bin/env python
"""Device time of the backbone tail hand-off against the eager PyTorch tail it replaces.
    python tools/tail_bench.py            one JSON line per (B, D, mode)
Algorithmic bytes, fp32: forward 8*B*D (read x, write p), backward 16*B*D (read p, dp, x; write dx); the BatchNorm passes
re-read from L1/L2.  [1024, 512] is 2 MiB per tensor: launch latency, not HBM, bounds every variant."""
import json
import os
import sys

R = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [R, os.path.join(R, 'very-large-scale-face-recognition_b200')]
import torch
import torch.nn as nn
import torch.nn.functional as F
import ffc_b200
from ffc_b200 import _capi

dev = torch.device('cuda')
peaks = json.load(open(os.path.join(R, 'MEASURED_PEAKS.json'))) if os.path.isfile(os.path.join(R, 'MEASURED_PEAKS.json')) else {'hbm_gbs': 6650.0}
lib = _capi.lib()


def timed(fn, iters=200, warm=20):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters * 1e3        # us


for B, D in [(1024, 512), (512, 512), (1024, 128), (8192, 512)]:
    for mode in ('bn_train', 'normalize'):
        x = torch.randn(B, D, device=dev, requires_grad=True)
        dp = torch.randn(B, D, device=dev)
        ours = ffc_b200.FFCTail(D).to(dev) if mode == 'bn_train' else ffc_b200.NormalizeTail()
        bn = nn.BatchNorm1d(D).to(dev)
        eager = (lambda t: F.normalize(bn(t))) if mode == 'bn_train' else F.normalize

        def step(f):
            x.grad = None
            f(x).backward(dp)
        n0 = lib.ffc_launch_count()
        step(ours)
        launches = lib.ffc_launch_count() - n0
        t_ours, t_eager = timed(lambda: step(ours)), timed(lambda: step(eager))
        # the kernels alone: the same launches through the bare C ABI (no autograd / allocator host time) -- still enqueued one by
        # one from Python, so this is an upper bound on the device time
        import ctypes as C
        xd, p, inv, st, dx = x.detach(), torch.empty(B, D, device=dev), torch.empty(B, device=dev), torch.empty(2, D, device=dev), torch.empty(B, D, device=dev)
        db = torch.empty(D, device=dev)
        ws = torch.zeros(8 << 20, dtype=torch.uint8, device=dev)
        m = ours if mode == 'bn_train' else None
        a = _capi.TailArgs(xd.data_ptr(), p.data_ptr(), D, inv.data_ptr(), B, D, 2 if m else 0, 1e-5, 0.1, m.weight.data_ptr() if m else None,
                           m.bias.data_ptr() if m else None, m.running_mean.data_ptr() if m else None, m.running_var.data_ptr() if m else None,
                           st[0].data_ptr() if m else None, st[1].data_ptr() if m else None, ws.data_ptr() if m else None, ws.numel() if m else 0)
        s0 = torch.cuda.current_stream().cuda_stream

        def raw():
            _capi.check(lib.ffc_tail_forward(C.byref(a), s0))
            _capi.check(lib.ffc_tail_backward(C.byref(a), dp.data_ptr(), D, dx.data_ptr(), None, db.data_ptr() if m else None, s0))
        t_raw = timed(raw, iters=1000)
        # device time: 50 forward + backward pairs captured in one CUDA graph (no host in the loop)
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g):
            s0 = torch.cuda.current_stream().cuda_stream
            for _ in range(50):
                raw()
        s0 = torch.cuda.current_stream().cuda_stream
        t_dev = timed(g.replay, iters=20, warm=3) / 50
        nbytes = 24 * B * D
        print(json.dumps(dict(B=B, D=D, mode=mode, launches_fwd_bwd=int(launches), us_ours=round(t_ours, 2), us_ours_c_abi=round(t_raw, 2), us_device=round(t_dev, 2), us_eager=round(t_eager, 2),
                              speedup=round(t_eager / t_ours, 2), algorithmic_bytes=nbytes, achieved_gbs=round(nbytes / (t_dev * 1e-6) / 1e9, 1),
                              peak_gbs=peaks['hbm_gbs'])))
