#!/usr/bin/env python
"""Whole training step (main.py:53-71: zero_grad, FFC.forward under fp16 autocast, scaled backward) with the REFERENCE's own backbones at
BASELINE.json's C1 / C2 model configurations, once with the reference's `FFC` module (its eager CUDA head: Python LRU loops, B x Q logits,
argsort) and once with `ffc_b200.FFC` as the drop-in -- same weights, same queue, same batches.  The backbones are outside the hot path
and identical in both arms; the difference is the head.  One JSON line per (config, arm).
    python tools/train_step_bench.py [c1] [c2]          (needs the staged reference tree oracle/_ref, see oracle/make_ref.py)"""
import contextlib
import io
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = '/root/reference' if os.path.isfile('/root/reference/ffc.py') else os.path.join(ROOT, 'oracle', '_ref')
sys.path[:0] = [ROOT, os.path.join(ROOT, 'very-large-scale-face-recognition_b200'), REF]

CONFIGS = {
    # mobilefacenet_def.py:78 / main.py:151-163 defaults; the reference MobileFaceNet turns NaN under cuDNN + fp16 autocast on this image
    # (tools/reference_mobilefacenet_nan_probe.py), so both arms run it on the native kernels
    'c1': dict(name='C1: MobileFaceNet + FFC head, 112x112, batch 64, 10k identities, queue 4096', net='mobile', D=128, B=64, Q=4096, N=10000,
               warm=4, steps=10, cudnn=False),
    'c2': dict(name='C2: iresnet50 (resnet_arcface) + FFC head, 112x112, batch 512, 100k identities, queue 65536', net='ir50', D=512, B=512, Q=65536,
               N=100000, warm=2, steps=5, cudnn=True),
}


def batches(c, n, seed):
    g = torch.Generator().manual_seed(seed)
    B, N, h = c['B'], c['N'], c['B'] // 2
    perm = torch.randperm(N, generator=torch.Generator().manual_seed(seed + 1))
    out = []
    for s in range(n):
        ids = perm[s * h:(s + 1) * h]                       # id half: the same identities in x and y (main.py:49-60)
        xl = torch.cat([ids, torch.randint(0, N, (B - h,), generator=g)])
        yl = torch.cat([ids, torch.randint(0, N, (B - h,), generator=g)])
        out.append((torch.randn(B, 3, 112, 112, generator=g), torch.randn(B, 3, 112, 112, generator=g), xl, yl))
    return out


def run(key):
    import ffc as ref_ffc
    import ffc_b200
    c = CONFIGS[key]
    dev = torch.device('cuda')
    torch.manual_seed(1)
    ref = ref_ffc.FFC(c['net'], c['D'], queue_size=c['Q'], scale=32.0, loss_type='Arc', margin=0.5).to(dev)
    ours = ffc_b200.FFC(c['net'], c['D'], queue_size=c['Q'], scale=32.0, loss_type='Arc', margin=0.5, max_batch=c['B'])
    ours.probe_net.load_state_dict(ref.probe_net.state_dict())
    ours.gallery_net.load_state_dict(ref.gallery_net.state_dict())
    ours.queue.copy_(ref.queue.detach().cpu())
    ours = ours.to(dev)
    data = [(x.to(dev), y.to(dev), xl, yl) for x, y, xl, yl in batches(c, c['warm'] + c['steps'], 7)]
    res = {}
    for arm, m in (('reference', ref), ('ours', ours)):
        m.train()
        losses = []
        with torch.backends.cudnn.flags(enabled=c['cudnn']):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            for s, (x, y, xl, yl) in enumerate(data):
                if s == c['warm']:
                    torch.cuda.synchronize()
                    e0.record()
                m.zero_grad(set_to_none=True)
                with contextlib.redirect_stdout(io.StringIO()), torch.amp.autocast('cuda'):      # main.py:64-65 (ffc.py:196 prints)
                    loss = m(x, y, xl, yl)
                (loss * 1024.0).backward()                                                        # main.py:69 (a GradScaler's factor)
                losses.append(loss.detach())
            e1.record()
            torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / c['steps']
        res[arm] = ms
        print(json.dumps(dict(metric='train_step_images_per_s', config=c['name'], impl=arm, value=2 * c['B'] / (ms * 1e-3), unit='images/s',
                              ms_per_step=ms, steps=c['steps'], warmup=c['warm'], autocast='fp16', cudnn=c['cudnn'],
                              last_loss=float(losses[-1]), data='synthetic 112x112 images, id half + instance half labels')), flush=True)
    print(json.dumps(dict(config=c['name'], step_time_ratio_reference_over_ours=res['reference'] / res['ours'])), flush=True)
    del ref, ours, data
    torch.cuda.empty_cache()


if __name__ == '__main__':
    for k in (sys.argv[1:] or ['c1', 'c2']):
        run(k)
