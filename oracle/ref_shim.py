"""Loader for the UNMODIFIED reference head (/root/reference/{ffc,lru}.py) as a CPU oracle.

TEST INFRASTRUCTURE ONLY.  In the build container it loads /root/reference -- it is what
tests/golden/make_golden.py uses to produce the committed fixtures, and what
tests/test_oracle_vs_reference.py uses for live cross-checks.  On the GPU box, where
/root/reference does not exist, it loads the byte-for-byte staging oracle/_ref/ written by
oracle/make_ref.py (git-ignored, shipped with the snapshot): bench.py's cpu_baseline /
--impl reference legs and the -m gpu tests that run the reference's own backbones.

No reference file is edited or copied; three process-local shims make the head
runnable on CPU:
  1. ``torch.Tensor.cuda`` -> identity            (ffc.py:179,180,194,237,238,246 call .cuda())
  2. ``ffc.create_net``    -> NormalizeNet        (embeddings in; ffc.py:22-23)
  3. ``allow_mutation_on_saved_tensors`` around fwd+bwd (ffc.py:182/241/255 write ``queue``
     in place after F.linear saved a view of it)
plus a spy on ``torch.LongTensor`` that records the integer bookkeeping in the
order ffc.py builds it (rows, cols, labels, ones) per pass.
"""
from __future__ import annotations

import contextlib
import importlib
import io
import os
import sys

import torch
import torch.nn.functional as F

_STAGED = os.path.join(os.path.dirname(os.path.abspath(__file__)), '_ref')     # oracle/make_ref.py: byte-for-byte staging for the GPU box


def _default_root():
    if os.path.isfile('/root/reference/ffc.py'):
        return '/root/reference'
    return _STAGED


REF_ROOT = os.environ.get('FFC_REFERENCE_ROOT') or _default_root()


def available() -> bool:
    return os.path.isfile(os.path.join(REF_ROOT, 'ffc.py'))


class NormalizeNet(torch.nn.Module):
    """Stand-in backbone: F.normalize only (every reference net ends in it)."""

    def __init__(self, *a, **k):
        super().__init__()
        self.dummy = torch.nn.Parameter(torch.zeros(1))

    def forward(self, x):
        return F.normalize(x + 0.0 * self.dummy)


def load():
    """Return (ffc_module, lru_module) imported from the reference tree."""
    assert available(), 'reference tree not present'
    if REF_ROOT not in sys.path:
        sys.path.insert(0, REF_ROOT)
    lru_mod = importlib.import_module('lru')
    ffc_mod = importlib.import_module('ffc')
    assert os.path.realpath(ffc_mod.__file__).startswith(os.path.realpath(REF_ROOT))
    ffc_mod.create_net = lambda *a, **k: NormalizeNet()
    return ffc_mod, lru_mod


@contextlib.contextmanager
def cpu_shims(spy=None):
    """Patch Tensor.cuda -> identity, silence prints, record LongTensor constructions."""
    orig_cuda = torch.Tensor.cuda
    orig_long = torch.LongTensor

    class SpyLong:
        def __new__(cls, data=()):
            if spy is not None:
                spy.append(list(data))
            return torch.tensor(list(data), dtype=torch.long)

    torch.Tensor.cuda = lambda self, *a, **k: self
    torch.LongTensor = SpyLong
    try:
        with contextlib.redirect_stdout(io.StringIO()):
            with torch.autograd.graph.allow_mutation_on_saved_tensors():
                yield
    finally:
        torch.Tensor.cuda = orig_cuda
        torch.LongTensor = orig_long


def make_ffc(feat_dim, queue_size, scale, loss_type, margin, queue=None):
    ffc_mod, _ = load()
    m = ffc_mod.FFC('x', feat_dim, queue_size=queue_size, scale=scale, loss_type=loss_type, margin=margin)
    if queue is not None:
        m.queue = queue.detach().clone().float()
    return m


def forward_backward(m, x, y, x_label, y_label):
    """One reference FFC.forward + backward on CPU.  Returns (loss, dx, dy, trace) where trace is
    [rollback{rows,cols,labels,ones}, commit{...}] and dx/dy are gradients w.r.t. the (unit-norm)
    inputs fed through NormalizeNet... taken w.r.t. the *embeddings* p, i.e. after normalisation."""
    spy = []
    grads = {}
    xs = x.detach().clone().float().requires_grad_(True)
    ys = y.detach().clone().float().requires_grad_(True)

    # capture the gradient w.r.t. the embeddings leaving the (normalising) backbone
    def hook_p(mod, inp, out):
        tag = 'x' if inp[0] is xs else 'y'
        if out.requires_grad:
            out.register_hook(lambda gr, tag=tag: grads.__setitem__(tag, gr.detach().clone()))

    h = m.probe_net.register_forward_hook(hook_p)
    try:
        with cpu_shims(spy):
            loss = m(xs, ys, torch.as_tensor(x_label, dtype=torch.long), torch.as_tensor(y_label, dtype=torch.long))
            if torch.is_tensor(loss) and loss.requires_grad:
                loss.backward()
    finally:
        h.remove()
    assert len(spy) == 8, len(spy)
    trace = []
    for i in (0, 4):
        trace.append(dict(rows=spy[i], cols=spy[i + 1], labels=spy[i + 2], ones=sorted(spy[i + 3])))
    B, D = x.shape
    dx = grads.get('x', torch.zeros(B, D))
    dy = grads.get('y', torch.zeros(B, D))
    return float(loss), dx, dy, trace
