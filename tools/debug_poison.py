#!/usr/bin/env python
"""debug: does any kernel read memory it did not write?  Poison the caching allocator's free blocks with NaN bit patterns, then run
head passes / the module and look for NaN."""
import contextlib, io, os, sys
R = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [R, os.path.join(R, 'very-large-scale-face-recognition_b200'), os.path.join(R, 'oracle', '_ref')]
import torch
import torch.nn.functional as F
import ffc_b200
from oracle.head_ref import HeadOracle
dev = torch.device('cuda')


def poison():
    ts = [torch.full((sz,), float('nan'), device=dev) for sz in (1 << 26, 1 << 22, 1 << 18, 1 << 14, 1 << 10, 64, 1)]
    ts += [torch.full((n,), float('nan'), device=dev) for n in (64 * 128, 64, 4 * 64, 3 * 64 * 3, 2 * 4096 * 128 // 2, 64 * 128 // 2) for _ in range(8)]
    torch.cuda.synchronize()
    del ts


torch.manual_seed(1)
D, Q, B = 128, 4096, 64
gen = torch.Generator().manual_seed(2)
perm = torch.randperm(3 * B, generator=gen)
h = B // 2
xl, yl = torch.cat([perm[:h], perm[h:B]]), torch.cat([perm[:h], perm[B:B + h]])
x = F.normalize(torch.randn(B, D, generator=gen)).to(dev)
y = F.normalize(torch.randn(B, D, generator=gen)).to(dev)
for prec in ('bf16', 'fp32'):
    poison()
    hd = ffc_b200.FFCHead(D, Q, 32.0, 'Arc', 0.5, precision=prec, max_batch=B, device=dev)
    q0 = hd.queue.detach().cpu().clone()
    o = HeadOracle(D, Q, 32.0, 'Arc', 0.5, queue=q0, dtype=torch.float64)
    for step in range(3):
        poison()
        l2, d2 = hd._pass(x, y, xl, yl, False)
        poison()
        l1, d1 = hd._pass(y, x, yl, xl, True)
        lo2 = o.head_pass(x.cpu().double(), y.cpu().double(), xl.tolist(), yl.tolist(), False)
        lo1 = o.head_pass(y.cpu().double(), x.cpu().double(), yl.tolist(), xl.tolist(), True)
        print(prec, step, 'ours rb/cm', float(l2), float(l1), 'oracle', float(lo2), float(lo1), 'dp finite', bool(torch.isfinite(d2).all()), bool(torch.isfinite(d1).all()),
              'nan rows rb', torch.nonzero(~torch.isfinite(d2).all(dim=1)).flatten().tolist()[:10], 'cm', torch.nonzero(~torch.isfinite(d1).all(dim=1)).flatten().tolist()[:10])
        poison()
        lp, dxp, dyp = hd.forward_pair(x, y, y, x, xl, yl)
        print('   forward_pair loss', float(lp), 'finite', bool(torch.isfinite(dxp).all()), bool(torch.isfinite(dyp).all()))
        o.forward(x.cpu().double(), y.cpu().double(), xl.tolist(), yl.tolist())
# the module with the reference backbone after a backward of another module (the failing test's sequence)
import ffc as ref_ffc
ref = ref_ffc.FFC('mobile', D, queue_size=Q, scale=32.0, loss_type='Arc', margin=0.5).to(dev)
ours = ffc_b200.FFC('mobile', D, queue_size=Q, scale=32.0, loss_type='Arc', margin=0.5, max_batch=B)
ours.probe_net.load_state_dict(ref.probe_net.state_dict())
ours.gallery_net.load_state_dict(ref.gallery_net.state_dict())
ours.queue.copy_(ref.queue.detach().cpu())
ours = ours.to(dev)
ref.train(), ours.train()
xi = torch.randn(B, 3, 112, 112, generator=gen).to(dev)
yi = torch.randn(B, 3, 112, 112, generator=gen).to(dev)
for m, name in ((ref, 'ref'), (ours, 'ours')):
    with contextlib.redirect_stdout(io.StringIO()), torch.amp.autocast('cuda'):
        loss = m(xi, yi, xl, yl)
    print(name, 'loss', float(loss))
    (loss * 1024.0).backward()
    gn = sum(float(p.grad.float().pow(2).sum()) for p in m.probe_net.parameters() if p.grad is not None) ** 0.5
    print(name, 'grad norm', gn)
# hooks on ours: which tensor is the first non-finite one?
ours.zero_grad(set_to_none=True)
with torch.amp.autocast('cuda'):
    p_rb = ours.probe_net(xi)
    ours._momentum_update_gallery()
    g_rb = ours.gallery_net(yi)
    p_cm = ours.probe_net(yi)
    g_cm = ours.gallery_net(xi)
for n, t in (('p_rb', p_rb), ('g_rb', g_rb), ('p_cm', p_cm), ('g_cm', g_cm)):
    print(n, t.dtype, 'finite', bool(torch.isfinite(t).all()))
print('gallery params finite', all(bool(torch.isfinite(p).all()) for p in ours.gallery_net.parameters()), 'probe', all(bool(torch.isfinite(p).all()) for p in ours.probe_net.parameters()))
print('gallery buffers finite', all(bool(torch.isfinite(b.float()).all()) for b in ours.gallery_net.buffers()))
