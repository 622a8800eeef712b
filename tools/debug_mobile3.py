#!/usr/bin/env python
"""debug: the failing sequences of tests/test_gpu_train.py with hooks that report the first non-finite leaf-module output"""
import contextlib, io, os, sys
R = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [R, os.path.join(R, 'very-large-scale-face-recognition_b200'), os.path.join(R, 'oracle', '_ref')]
import torch
import ffc as ref_ffc
import ffc_b200
dev = torch.device('cuda')


def fin(t):
    return bool(torch.isfinite(t.float()).all())


def watch(net, tag, log):
    hs = []
    if os.environ.get('NOHOOK'):
        return hs
    for name, mod in net.named_modules():
        if len(list(mod.children())) == 0:
            def hook(m, i, o, name=name):
                if len(log) < 4 and not fin(o):
                    st = {k: fin(v) for k, v in list(m.named_parameters(recurse=False)) + list(m.named_buffers(recurse=False))}
                    log.append((tag, name, type(m).__name__, 'in finite', fin(i[0]), 'in absmax', float(i[0].float().abs().max()), 'in dtype', str(i[0].dtype), 'state finite', st))
            hs.append(mod.register_forward_hook(hook))
    return hs


mode = sys.argv[1] if len(sys.argv) > 1 else 'dropin'
torch.manual_seed(1 if mode == 'dropin' else 0)
D, Q, B = 128, 4096, 64
log = []
if mode == 'dropin':
    ref = ref_ffc.FFC('mobile', D, queue_size=Q, scale=32.0, loss_type='Arc', margin=0.5).to(dev)
    ours = ffc_b200.FFC('mobile', D, queue_size=Q, scale=32.0, loss_type='Arc', margin=0.5, max_batch=B)
    ours.probe_net.load_state_dict(ref.probe_net.state_dict())
    ours.gallery_net.load_state_dict(ref.gallery_net.state_dict())
    ours.queue.copy_(ref.queue.detach().cpu())
    ours = ours.to(dev)
    ref.train(), ours.train()
    watch(ours.probe_net, 'ours.probe', log), watch(ours.gallery_net, 'ours.gallery', log), watch(ref.probe_net, 'ref.probe', log)
    gen = torch.Generator().manual_seed(2)
    for s in range(3):
        perm = torch.randperm(3 * B, generator=gen)
        h = B // 2
        xl, yl = torch.cat([perm[:h], perm[h:B]]), torch.cat([perm[:h], perm[B:B + h]])
        x = torch.randn(B, 3, 112, 112, generator=gen).to(dev)
        y = torch.randn(B, 3, 112, 112, generator=gen).to(dev)
        for m, name in ((ref, 'ref'), (ours, 'ours')):
            m.zero_grad(set_to_none=True)
            with contextlib.redirect_stdout(io.StringIO()), torch.amp.autocast('cuda'):
                loss = m(x, y, xl, yl)
            (loss * 1024.0).backward()
            print(s, name, 'loss', float(loss), 'input finite', fin(x), fin(y), flush=True)
else:
    from ffc_b200 import train as T
    net = ffc_b200.FFC('mobile', 128, queue_size=4096, scale=32.0, loss_type='Arc', margin=0.5, max_batch=64, precision=os.environ.get('PREC', 'bf16')).to(dev)
    if os.environ.get('EAGER_EMA'):
        def eager_ema(self=net):
            with torch.no_grad():
                for pp, pg in zip(self.probe_net.parameters(), self.gallery_net.parameters()):
                    pg.data = pg.data * self.m + pp.data * (1. - self.m)
        net._momentum_update_gallery = eager_ema
    if os.environ.get('NOPREFETCH'):
        net.prefetch_labels = None
    if os.environ.get('STASH'):
        from ffc_b200.ffc import _HeadPairFn
        stash = {}

        def fwd(x, y, x_label, y_label, self=net):
            p_rb = self.probe_net(x)
            with torch.no_grad():
                self._momentum_update_gallery()
                g_rb = self.gallery_net(y)
            p_cm = self.probe_net(y)
            with torch.no_grad():
                g_cm = self.gallery_net(x)
            loss = _HeadPairFn.apply(p_rb, p_cm, self, g_rb, g_cm, x_label, y_label)
            stash.update(x=x, y=y, p_rb=p_rb.detach(), g_rb=g_rb, p_cm=p_cm.detach(), g_cm=g_cm, loss=loss.detach())
            return loss
        net.forward = fwd
    watch(net.probe_net, 'probe', log), watch(net.gallery_net, 'gallery', log)
    opt = torch.optim.SGD([p for p in net.parameters() if p.requires_grad], lr=0.05, momentum=0.9, weight_decay=1e-4, nesterov=True)
    scaler = torch.amp.GradScaler('cuda')
    src = T.SyntheticSource(num_class=10000, batch_size=64, image_size=112, batches_per_epoch=8, seed=3)
    logs = []
    def logit(d):
        logs.append(d)
        if os.environ.get('STASH'):
            print(d['iter'], {k: (fin(v), str(v.dtype)) for k, v in stash.items()}, 'queue finite', fin(net.queue), 'bn buffers finite',
                  all(fin(b) for b in net.probe_net.buffers()), all(fin(b) for b in net.gallery_net.buffers()), flush=True)
    T.train_one_epoch(src.id_loader(), src.instance_loader(), net, opt, scaler, save_every=int(os.environ.get('SAVE_EVERY', 2)), device=dev, log=logit)
    print([round(l['loss'], 3) for l in logs], 'scale', scaler.get_scale())
    print('params finite', all(fin(p) for p in net.probe_net.parameters()), 'gallery params finite', all(fin(p) for p in net.gallery_net.parameters()))
for l in log:
    print(l)
