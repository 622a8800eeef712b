#!/usr/bin/env python
"""debug: does the REFERENCE module alone (its MobileFaceNet under fp16 autocast, scaled backward) go NaN at the second iteration on this stack?"""
import contextlib, io, os, sys
R = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [R, os.path.join(R, 'oracle', '_ref')]
import torch
import ffc as ref_ffc
from model import create_net
dev = torch.device('cuda')
if os.environ.get('NOCUDNN'):
    torch.backends.cudnn.enabled = False
torch.manual_seed(0)
B = 64
mode = sys.argv[1] if len(sys.argv) > 1 else 'ffc'
gen = torch.Generator().manual_seed(3)
if mode == 'ffc':
    m = ref_ffc.FFC('mobile', 128, queue_size=4096, scale=32.0, loss_type='Arc', margin=0.5).to(dev)
    m.train()
    for it in range(4):
        x = torch.randn(B, 3, 112, 112, generator=gen).to(dev)
        y = torch.randn(B, 3, 112, 112, generator=gen).to(dev)
        perm = torch.randperm(10000, generator=gen)
        xl, yl = perm[:B], torch.cat([perm[:B // 2], perm[B:B + B // 2]])
        m.zero_grad(set_to_none=True)
        with contextlib.redirect_stdout(io.StringIO()), torch.amp.autocast('cuda'):
            loss = m(x, y, xl, yl)
        (loss * 65536.0).backward()
        gn = [bool(torch.isfinite(p.grad).all()) for p in m.probe_net.parameters() if p.grad is not None]
        print(it, 'reference FFC loss', float(loss), 'grads finite', all(gn), flush=True)
else:
    net = create_net('mobile', feat_dim=128, fp16=True).to(dev).train()
    dt = torch.bfloat16 if mode == 'bf16' else torch.float16
    for it in range(4):
        x = torch.randn(B, 3, 112, 112, generator=gen).to(dev)
        with torch.amp.autocast('cuda', dtype=dt):
            p = net(x)
            loss = (p * torch.randn(B, 128, device=dev)).sum()
        net.zero_grad(set_to_none=True)
        (loss * 65536.0 * 64).backward()
        gn = [bool(torch.isfinite(q.grad).all()) for q in net.parameters() if q.grad is not None]
        print(it, mode, 'backbone only: output finite', bool(torch.isfinite(p).all()), 'grads finite', all(gn), flush=True)
