"""Golden vectors for the backbone tail from the UNMODIFIED reference backbones on CPU.

Run in the build container only (the reference does not travel to the GPU box):
    python tests/golden/make_golden_tail.py
A real forward + backward of ``model.resnet_arcface.iresnet50`` and ``model.mobilefacenet_def.MobileFaceNet`` on seeded
112x112 images; forward hooks capture what enters the tail (the ``fc`` / flattened ``linear1`` output) and autograd gives
its gradient.  Writes tests/golden/tail_<case>.npz (committed).  No reference source is copied.
"""
import os
import sys

import numpy as np
import torch

REF = os.environ.get('FFC_REFERENCE_ROOT', '/root/reference')
OUT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, REF)


def _capture(cap):
    def hook(module, inputs, output):        # returns None: the output is observed, not replaced
        output.retain_grad()
        cap['x'] = output
    return hook


def run_iresnet(case, B, feat_dim, training, steps):
    from model.resnet_arcface import iresnet50
    torch.manual_seed(11)
    net = iresnet50(feat_dim=feat_dim, fp16=False)
    with torch.no_grad():                                   # non-trivial affine / running state (init is bias 0, mean 0, var 1)
        net.features.bias.normal_(0, 0.3)
        net.features.running_mean.normal_(0, 0.2)
        net.features.running_var.uniform_(0.5, 1.5)
    if not training:
        # an untrained iresnet overflows fp32 in eval() (identity running statistics through 50 layers): calibrate the running
        # statistics of every BatchNorm on one batch first, then perturb the tail's again so they differ from any batch's
        moms = {m: m.momentum for m in net.modules() if isinstance(m, torch.nn.modules.batchnorm._BatchNorm)}
        for m in moms:
            m.momentum = 1.0
        net.train()
        with torch.no_grad():
            net(torch.randn(8, 3, 112, 112))
            net.features.running_mean.add_(0.2 * torch.randn(feat_dim))
            net.features.running_var.mul_(torch.empty(feat_dim).uniform_(0.7, 1.4))
        for m, v in moms.items():
            m.momentum = v
    net.train(training)
    rec = {}
    for s in range(steps):
        cap = {}
        h = net.fc.register_forward_hook(_capture(cap))
        rm0, rv0 = net.features.running_mean.clone(), net.features.running_var.clone()
        img = torch.randn(B, 3, 112, 112)
        p = net(img)
        dp = torch.randn(B, feat_dim)
        net.zero_grad()
        (p * dp).sum().backward()
        h.remove()
        rec.update({f'x{s}': cap['x'].detach().numpy(), f'p{s}': p.detach().numpy(), f'dp{s}': dp.numpy(), f'dx{s}': cap['x'].grad.numpy(),
                    f'dbias{s}': net.features.bias.grad.numpy().copy(), f'rm_in{s}': rm0.numpy(), f'rv_in{s}': rv0.numpy(),
                    f'rm_out{s}': net.features.running_mean.numpy().copy(), f'rv_out{s}': net.features.running_var.numpy().copy()})
    rec.update(weight=net.features.weight.detach().numpy(), bias=net.features.bias.detach().numpy(), eps=np.float64(net.features.eps),
               momentum=np.float64(net.features.momentum), training=np.bool_(training), steps=np.int64(steps), bn=np.bool_(True),
               weight_requires_grad=np.bool_(net.features.weight.requires_grad))
    np.savez_compressed(os.path.join(OUT, f'tail_{case}.npz'), **rec)
    print(case, 'ok', {k: v.shape for k, v in rec.items() if k.endswith('0')})


def run_mobile(case, B, feat_dim):
    from model.mobilefacenet_def import MobileFaceNet
    torch.manual_seed(12)
    net = MobileFaceNet(feat_dim=feat_dim, fp16=False)
    net.train()
    cap = {}
    h = net.linear1.register_forward_hook(_capture(cap))
    img = torch.randn(B, 3, 112, 112)
    p = net(img)
    dp = torch.randn(B, feat_dim)
    (p * dp).sum().backward()
    h.remove()
    x = cap['x']
    rec = dict(x0=x.detach().flatten(1).numpy(), p0=p.detach().numpy(), dp0=dp.numpy(), dx0=x.grad.flatten(1).numpy(), bn=np.bool_(False),
               steps=np.int64(1))
    np.savez_compressed(os.path.join(OUT, f'tail_{case}.npz'), **rec)
    print(case, 'ok', {k: v.shape for k, v in rec.items() if k.endswith('0')})


if __name__ == '__main__':
    run_iresnet('ir50_train', B=6, feat_dim=512, training=True, steps=2)
    run_iresnet('ir50_eval', B=5, feat_dim=512, training=False, steps=1)
    run_iresnet('ir50_d128_train', B=4, feat_dim=128, training=True, steps=1)
    run_mobile('mobile_d128', B=6, feat_dim=128)
