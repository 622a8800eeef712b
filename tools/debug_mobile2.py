#!/usr/bin/env python
"""debug: where does the NaN of ours.probe_net (reference MobileFaceNet under fp16 autocast) come from?"""
import contextlib, io, os, sys
R = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [R, os.path.join(R, 'very-large-scale-face-recognition_b200'), os.path.join(R, 'oracle', '_ref')]
import torch
import ffc as ref_ffc
import ffc_b200
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
xi = torch.randn(B, 3, 112, 112, generator=gen).to(dev)
yi = torch.randn(B, 3, 112, 112, generator=gen).to(dev)


def fin(t):
    return bool(torch.isfinite(t.float()).all())


def probe(net, tag):
    with torch.amp.autocast('cuda'):
        out = net(xi)
    print(f'{tag}: output finite {fin(out)}', flush=True)
    return out


def same_weights(a, b):
    sa, sb = a.state_dict(), b.state_dict()
    return all(torch.equal(sa[k], sb[k]) for k in sa)


print('weights equal before anything:', same_weights(ours.probe_net, ref.probe_net))
probe(ours.probe_net, 'ours.probe_net first call')
probe(ref.probe_net, 'ref.probe_net first call')
with contextlib.redirect_stdout(io.StringIO()), torch.amp.autocast('cuda'):
    loss = ref(xi, yi, xl, yl)
print('ref module loss', float(loss), flush=True)
(loss * 1024.0).backward()
print('after ref backward: weights equal', same_weights(ours.probe_net, ref.probe_net))
probe(ours.probe_net, 'ours.probe_net after ref fwd+bwd')
with torch.amp.autocast('cuda'):
    lo = ours(xi, yi, xl, yl)
print('ours module loss', float(lo), flush=True)
probe(ours.probe_net, 'ours.probe_net after ours fwd')
print('params finite', all(fin(p) for p in ours.probe_net.parameters()), 'buffers finite', all(fin(b) for b in ours.probe_net.buffers()))
bad = [n for n, b in ours.probe_net.named_buffers() if not fin(b)]
print('non-finite buffers:', bad[:8])
diff = [k for k in ours.probe_net.state_dict() if not torch.equal(ours.probe_net.state_dict()[k], ref.probe_net.state_dict()[k]) and 'running' not in k and 'num_batches' not in k]
print('parameters that differ from ref now:', diff[:8], len(diff))
# first non-finite module output
first = []
hooks = []
for name, mod in ours.probe_net.named_modules():
    if len(list(mod.children())) == 0:
        hooks.append(mod.register_forward_hook(lambda m, i, o, name=name: first.append((name, type(m).__name__, fin(i[0]), fin(o))) if not fin(o) and len(first) < 3 else None))
probe(ours.probe_net, 'ours.probe_net with hooks')
print('first non-finite leaf modules:', first)
for hk in hooks:
    hk.remove()
ours.probe_net.eval()
probe(ours.probe_net, 'ours.probe_net eval mode')
