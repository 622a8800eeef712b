#!/usr/bin/env python
"""debug: NaN loss of ffc_b200.FFC with the reference MobileFaceNet (tests/test_gpu_train.py)"""
import contextlib, io, os, sys
R = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [R, os.path.join(R, 'very-large-scale-face-recognition_b200'), os.path.join(R, 'oracle', '_ref')]
import torch
import ffc as ref_ffc
import ffc_b200
from oracle.head_ref import HeadOracle
dev = torch.device('cuda')
torch.manual_seed(1)
D, Q, B = 128, 4096, 64
ref = ref_ffc.FFC('mobile', D, queue_size=Q, scale=32.0, loss_type='Arc', margin=0.5).to(dev)
ours = ffc_b200.FFC('mobile', D, queue_size=Q, scale=32.0, loss_type='Arc', margin=0.5, max_batch=B)
ours.probe_net.load_state_dict(ref.probe_net.state_dict())
ours.gallery_net.load_state_dict(ref.gallery_net.state_dict())
ours.queue.copy_(ref.queue.detach().cpu())
ours = ours.to(dev)
ref.train(), ours.train()
gen = torch.Generator().manual_seed(2)
perm = torch.randperm(3 * B, generator=gen)
h = B // 2
xl, yl = torch.cat([perm[:h], perm[h:B]]), torch.cat([perm[:h], perm[B:B + h]])
x = torch.randn(B, 3, 112, 112, generator=gen).to(dev)
y = torch.randn(B, 3, 112, 112, generator=gen).to(dev)
with torch.amp.autocast('cuda'):
    p_rb = ours.probe_net(x)
    g_rb = ours.gallery_net(y)
    p_cm = ours.probe_net(y)
    g_cm = ours.gallery_net(x)
for n, t in (('p_rb', p_rb), ('g_rb', g_rb), ('p_cm', p_cm), ('g_cm', g_cm)):
    print(n, t.dtype, tuple(t.shape), 'finite', bool(torch.isfinite(t).all()), 'norm min/max', float(t.float().norm(dim=1).min()), float(t.float().norm(dim=1).max()),
          'contig', t.is_contiguous())
for prec in ('bf16', 'fp32'):
    hd = ffc_b200.FFCHead(D, Q, 32.0, 'Arc', 0.5, precision=prec, max_batch=B, device=dev)
    hd.queue.copy_(ours.queue)
    hd._ensure(); hd.sync_mirror()
    o = HeadOracle(D, Q, 32.0, 'Arc', 0.5, queue=ours.queue.cpu(), dtype=torch.float64)
    l2, d2 = hd._pass(p_rb.detach().float(), g_rb.detach().float(), xl, yl, False)
    bk = hd.last_bookkeeping()
    l1, d1 = hd._pass(p_cm.detach().float(), g_cm.detach().float(), yl, xl, True)
    lo2 = o.head_pass(p_rb.detach().cpu().double(), g_rb.detach().cpu().double(), xl.tolist(), yl.tolist(), False)
    lo1 = o.head_pass(p_cm.detach().cpu().double(), g_cm.detach().cpu().double(), yl.tolist(), xl.tolist(), True)
    print(prec, 'ours rb/cm', float(l2), float(l1), 'oracle', float(lo2), float(lo1), 'dp finite', bool(torch.isfinite(d2).all()), bool(torch.isfinite(d1).all()))
    print('  labels rb', bk[2][:8], '... n_out', sum(v < 0 for v in bk[2]), 'row_loss nan rows:', 'n/a')
    tg = hd.stats['tgt'] if B == hd.stats['tgt'].shape[1] else hd._compact[('tgt', B)]
    print('  tgt cos range', float(tg[0].min()), float(tg[0].max()), float(tg[1].min()), float(tg[1].max()))
with contextlib.redirect_stdout(io.StringIO()), torch.amp.autocast('cuda'):
    lr = ref(x, y, xl, yl)
    lo = ours(x, y, xl, yl)
print('module losses ref / ours', float(lr), float(lo))
