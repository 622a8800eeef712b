"""Backbone tail -> head hand-off: the last two operations of every reference backbone,
``F.normalize(self.features(x))`` with ``features = nn.BatchNorm1d(feat_dim, eps=1e-05)`` (resnet_arcface.py:99,151;
resnet_std.py:201-202) or plain ``F.normalize(x)`` (mobilefacenet_def.py:113-114), forward and backward, on the
`ffc_tail_*` kernels of libffc_b200 (csrc/tail.cu): one or two launches each way instead of ~6 + ~12 eager ones, and the
unit-norm rows can be written straight into the buffer the head reads (``out=``).

:class:`FFCTail` subclasses ``nn.BatchNorm1d``: same constructor, parameters, buffers and ``state_dict`` keys, same
``train()`` / ``eval()`` meaning, so it replaces ``features`` in place (:func:`fuse_tail`) and checkpoints written by the
reference load unchanged.  There is no CPU fallback.
"""
from __future__ import annotations

import ctypes as C

import torch
import torch.nn as nn

from . import _capi
from ._capi import TailArgs, check, ptr

MODES = {'normalize': 0, 'bn_eval': 1, 'bn_train': 2}


_WORKSPACES = {}


def _workspace(dev, stream, B, D):
    """Scratch of the BatchNorm kernels (slab partials + arrival counters): zero-filled once, left reusable by every call.  One per
    (device, stream): calls on one stream are ordered, calls on different streams must not share it."""
    need = C.c_int64()
    check(_capi.lib().ffc_tail_workspace_bytes(B, D, C.byref(need)))
    key = (dev.index, stream)
    ws = _WORKSPACES.get(key)
    if ws is None or ws.numel() < need.value:
        ws = torch.zeros(max(need.value, 1 << 20), dtype=torch.uint8, device=dev)      # filled on dev's current stream == `stream`
        _WORKSPACES[key] = ws
    return ws


def _f32c(t):
    return None if t is None else t.detach().to(dtype=torch.float32).contiguous()


class _TailFn(torch.autograd.Function):
    """p = normalize(batchnorm(x)); saves x, p, 1/||y|| and the statistics used."""

    @staticmethod
    def forward(ctx, x, weight, bias, running_mean, running_var, mode, eps, momentum, out_box):
        out = out_box[0]          # boxed: the destination is plain memory to autograd, not an input that is modified in place
        if not x.is_cuda:
            raise _capi.FFCError('the FFC tail runs on a CUDA device only (no CPU fallback)')
        if x.dim() != 2:
            raise ValueError(f'expected 2D input [batch, feat_dim] (got {x.dim()}D input)')      # nn.BatchNorm1d._check_input_dim, 2-D case
        B, D = x.shape
        if mode == MODES['bn_train'] and B <= 1:
            raise ValueError(f'Expected more than 1 value per channel when training, got input size {x.size()}')   # torch.nn.functional.batch_norm
        lib, dev = _capi.lib(), x.device
        x32 = _f32c(x)
        if out is None:
            out = torch.empty(B, D, dtype=torch.float32, device=dev)
        else:
            assert out.is_cuda and out.dtype == torch.float32 and out.shape == (B, D) and out.stride(1) == 1 and out.stride(0) >= D, \
                'out= must be a CUDA fp32 [B, D] view with unit column stride'
            assert not out.requires_grad, 'out= must not require grad'
            out = out.detach()          # a fresh alias of the destination: the result gets its own autograd identity
        w32, b32 = _f32c(weight), _f32c(bias)
        bn = mode != MODES['normalize']
        inv_norm = torch.empty(B, dtype=torch.float32, device=dev)
        stats = torch.empty(2, D, dtype=torch.float32, device=dev) if bn else None
        stream = torch.cuda.current_stream(dev).cuda_stream
        ws = _workspace(dev, stream, B, D) if bn else None
        args = TailArgs(x32.data_ptr(), out.data_ptr(), out.stride(0), inv_norm.data_ptr(), B, D, mode, float(eps), float(momentum),
                        ptr(w32), ptr(b32), ptr(running_mean), ptr(running_var),
                        stats[0].data_ptr() if bn else None, stats[1].data_ptr() if bn else None, ptr(ws), ws.numel() if bn else 0)
        with torch.cuda.device(dev):
            check(lib.ffc_tail_forward(C.byref(args), stream))
        ctx.save_for_backward(x32)
        # everything else `args` points into.  `out` is read by backward through its address: it may live in a staging buffer
        # whose other columns are legitimately written in between (torch's version counter is per storage)
        # (a detached alias, not `out` itself: `out` is this node's output, holding it here would be a reference cycle that only
        # the cyclic GC frees)
        ctx.args, ctx.keep = args, (out.detach(), inv_norm, stats, w32, b32, running_mean, running_var)
        ctx.in_dtype = x.dtype
        ctx.grads = (weight is not None and weight.requires_grad, bias is not None and bias.requires_grad)
        ctx.param_dtypes = (None if weight is None else weight.dtype, None if bias is None else bias.dtype)
        return out

    @staticmethod
    def backward(ctx, dp):
        (x32,) = ctx.saved_tensors
        B, D = x32.shape
        dev = x32.device
        dp = dp.to(torch.float32)
        if dp.stride(1) != 1 or dp.stride(0) < D:
            dp = dp.contiguous()
        dx = torch.empty(B, D, dtype=torch.float32, device=dev)
        want_w, want_b = ctx.grads
        dw = torch.empty(D, dtype=torch.float32, device=dev) if want_w else None
        db = torch.empty(D, dtype=torch.float32, device=dev) if want_b else None
        stream = torch.cuda.current_stream(dev).cuda_stream
        if ctx.args.workspace:                      # backward may run on another stream than forward did: take that stream's scratch
            ws = _workspace(dev, stream, B, D)
            ctx.args.workspace, ctx.args.workspace_bytes = ws.data_ptr(), ws.numel()
        with torch.cuda.device(dev):
            check(_capi.lib().ffc_tail_backward(C.byref(ctx.args), dp.data_ptr(), dp.stride(0), dx.data_ptr(), ptr(dw), ptr(db), stream))
        if dw is not None:
            dw = dw.to(ctx.param_dtypes[0])
        if db is not None:
            db = db.to(ctx.param_dtypes[1])
        return dx.to(ctx.in_dtype), dw, db, None, None, None, None, None, None


def l2_normalize(x, out=None):
    """``F.normalize(x)`` for 2-D x (mobilefacenet_def.py:113-114) on the tail kernels; differentiable."""
    return _TailFn.apply(x, None, None, None, None, MODES['normalize'], 0.0, 0.0, (out,))


class FFCTail(nn.BatchNorm1d):
    """``F.normalize(nn.BatchNorm1d(...)(x))`` as one op.  Everything about the module (arguments, parameters, buffers,
    state_dict, train / eval, ``momentum=None`` cumulative averaging, ``track_running_stats=False``) is nn.BatchNorm1d's;
    only 2-D input [batch, feat_dim] is accepted, which is what the reference backbones feed it."""

    def forward(self, x, out=None):
        if x.dim() != 2 or x.shape[1] != self.num_features:
            raise ValueError(f'FFCTail expects [batch, {self.num_features}] input, got {tuple(x.shape)}')
        if not x.is_cuda:       # before any state (num_batches_tracked) is touched
            raise _capi.FFCError('the FFC tail runs on a CUDA device only (no CPU fallback)')
        if self.training and x.shape[0] <= 1:
            raise ValueError(f'Expected more than 1 value per channel when training, got input size {x.size()}')
        use_batch = self.training or self.running_mean is None                   # nn.modules.batchnorm._BatchNorm.forward: bn_training
        momentum = 0.0 if self.momentum is None else self.momentum
        if self.training and self.track_running_stats and self.num_batches_tracked is not None:
            self.num_batches_tracked.add_(1)
            if self.momentum is None:
                momentum = 1.0 / float(self.num_batches_tracked)
        track = self.training and self.track_running_stats                          # eval-mode batch statistics never touch the buffers
        rm = self.running_mean if (track or not use_batch) else None
        rv = self.running_var if (track or not use_batch) else None
        for t in (rm, rv):
            if t is not None and not (t.is_cuda and t.dtype == torch.float32 and t.is_contiguous()):
                raise _capi.FFCError('FFCTail: running statistics must be contiguous fp32 CUDA buffers (no CPU fallback)')
        mode = MODES['bn_train'] if use_batch else MODES['bn_eval']
        return _TailFn.apply(x, self.weight, self.bias, rm, rv, mode, self.eps, momentum, (out,))

    @classmethod
    def from_batchnorm(cls, bn: nn.BatchNorm1d):
        """A tail that SHARES the parameters and buffers of an existing BatchNorm1d (so optimisers, EMA loops and state_dicts that
        already hold them keep working)."""
        t = cls(bn.num_features, eps=bn.eps, momentum=bn.momentum, affine=bn.affine, track_running_stats=bn.track_running_stats)
        for name in ('weight', 'bias'):
            t._parameters[name] = bn._parameters.get(name)
        for name in ('running_mean', 'running_var', 'num_batches_tracked'):
            t._buffers[name] = bn._buffers.get(name)
        t.train(bn.training)
        return t


class NormalizeTail(nn.Module):
    """``F.normalize`` as a module (MobileFaceNet's tail)."""

    def forward(self, x, out=None):
        return l2_normalize(x, out)


def fuse_tail(net: nn.Module, attr: str = 'features'):
    """Replace ``net.<attr>`` (the reference backbones' final ``nn.BatchNorm1d``, resnet_arcface.py:99 / resnet_std.py) by an
    :class:`FFCTail` sharing its parameters.  The backbone's own ``F.normalize`` that follows (resnet_arcface.py:151) then sees
    unit-norm rows: drop it from ``forward`` or leave it (idempotent up to 1 ulp, at the price of its eager launches)."""
    bn = getattr(net, attr)
    if isinstance(bn, FFCTail):
        return net
    if not isinstance(bn, nn.BatchNorm1d):
        raise TypeError(f'{type(net).__name__}.{attr} is {type(bn).__name__}, not nn.BatchNorm1d')
    setattr(net, attr, FFCTail.from_batchnorm(bn))
    return net
