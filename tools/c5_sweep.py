#!/usr/bin/env python
"""BASELINE.json config C5: FFC head sweep over emb dim 128/256/512 x queue 16k..2M with ArcFace vs CosFace margin, per-shape
roofline report (one GPU, B = 1024 probe rows per pass, all identities resident, synthetic unit embeddings).
    python tools/c5_sweep.py > gpurun_out/c5.md"""
import json
import os
import sys

R = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [R, os.path.join(R, 'very-large-scale-face-recognition_b200')]
import torch
import torch.nn.functional as F
import ffc_b200

dev = torch.device('cuda')
pk = json.load(open(os.path.join(R, 'MEASURED_PEAKS.json'))) if os.path.isfile(os.path.join(R, 'MEASURED_PEAKS.json')) else {'bf16_tflops': 1590.0}
B, steps, warm = 1024, 8, 2
print('| D | queue | loss | sweep ms | TFLOP/s (4BQD) | of burst bf16 peak | step ms (2 passes) | samples/s |')
print('|---:|---:|---|---:|---:|---:|---:|---:|')
DS = [int(v) for v in os.environ.get('C5_D', '128,256,512').split(',')]
QS = [int(v) for v in os.environ.get('C5_Q', '16384,65536,262144,1048576,2097152').split(',')]
for D in DS:
    for Q in QS:
        for loss_type, margin in ((('Arc', 0.5), ('AM', 0.4)) if not os.environ.get('C5_ARC_ONLY') else (('Arc', 0.5),)):
            torch.manual_seed(0)
            h = ffc_b200.FFCHead(D, Q, 32.0, loss_type, margin, precision='bf16', max_batch=B, device=dev)
            h._ensure()
            h.lru.restore_arrays(torch.arange(Q, dtype=torch.int64), torch.arange(Q, dtype=torch.int32))
            gen = torch.Generator().manual_seed(1)
            data = []
            for s in range(steps + warm):
                ids = torch.randperm(Q, generator=gen)[:B // 2]
                xl = torch.cat([ids, torch.randint(0, Q, (B - B // 2,), generator=gen)]).to(dev)
                yl = torch.cat([ids, torch.randint(0, Q, (B - B // 2,), generator=gen)]).to(dev)
                data.append((F.normalize(torch.randn(B, D, generator=gen)).to(dev), F.normalize(torch.randn(B, D, generator=gen)).to(dev), xl, yl))
            for s in range(warm):
                h.forward_pair(data[s][0], data[s][1], data[s][1], data[s][0], data[s][2], data[s][3])
            torch.cuda.synchronize()
            h.set_timing(True)
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for s in range(warm, warm + steps):
                h.forward_pair(data[s][0], data[s][1], data[s][1], data[s][0], data[s][2], data[s][3])
            e1.record()
            torch.cuda.synchronize()
            ms = e0.elapsed_time(e1) / steps
            sw_ms, sw_n = h.get_timing()
            h.set_timing(False)
            t = sw_ms / max(1, sw_n)
            tf = 4.0 * B * Q * D / (t * 1e-3) / 1e12
            print(f'| {D} | {Q} | {loss_type} m={margin} | {t:.3f} | {tf:.0f} | {tf / pk["bf16_tflops"]:.2f} | {ms:.3f} | {2 * B / (ms * 1e-3):.0f} |', flush=True)
            del h, data
            torch.cuda.empty_cache()
