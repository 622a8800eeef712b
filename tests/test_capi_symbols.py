"""CPU-side checks of the C-ABI boundary: the library loads and exports every symbol include/ffc_b200.h declares."""
import ctypes
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared():
    src = open(os.path.join(ROOT, 'include', 'ffc_b200.h')).read()
    src = re.sub(r'/\*.*?\*/', '', src, flags=re.S)
    return sorted(set(re.findall(r'\b(ffc_[a-z0-9_]+)\s*\(', src)))


def test_header_and_binding_agree():
    from ffc_b200 import _capi
    assert sorted(_capi.PROTOTYPES) == _declared()


def test_library_exports_every_declared_symbol():
    from ffc_b200 import _capi
    import importlib.util
    spec = importlib.util.spec_from_file_location('ffc_build', os.path.join(ROOT, 'very-large-scale-face-recognition_b200', 'build.py'))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    mod.build()
    lib = ctypes.CDLL(_capi.LIB_PATH)
    for name in _declared():
        assert hasattr(lib, name), name
    assert b'sm_100a' in _capi.lib().ffc_version()
    assert _capi.lib().ffc_launch_count() == 0


def test_no_cpu_fallback():
    import torch
    import ffc_b200
    if torch.cuda.is_available():
        pytest.skip('GPU present')
    with pytest.raises(ffc_b200.FFCError):
        ffc_b200.LRU(4)
    m = ffc_b200.FFC('identity', 16, queue_size=32)
    x = torch.nn.functional.normalize(torch.randn(4, 16))
    with pytest.raises(ffc_b200.FFCError):
        m(x, x, torch.arange(4), torch.arange(4))


def test_product_never_imports_oracle():
    pkg = os.path.join(ROOT, 'very-large-scale-face-recognition_b200')
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith(('.py', '.cu', '.cuh', '.h')):
                txt = open(os.path.join(dirpath, f)).read()
                assert 'oracle' not in txt.replace('oracle/', 'ORACLE_DOC'), os.path.join(dirpath, f)


def test_tail_argument_checks_run_on_the_host():
    """ffc_tail_* validate their arguments before touching the device: the error paths are testable without a GPU"""
    import ctypes as C
    from ffc_b200 import _capi
    lib = _capi.lib()
    launches = lib.ffc_launch_count()
    need = C.c_int64()
    assert lib.ffc_tail_workspace_bytes(1024, 512, C.byref(need)) == 0
    # 16 column-group counters (rounded to 256 B) + max(32 forward slabs, 64 backward CTAs) x 2 x D floats
    assert need.value == 256 + 64 * 2 * 512 * 4
    assert lib.ffc_tail_workspace_bytes(8192, 512, C.byref(need)) == 0 and need.value == 256 + 128 * 2 * 512 * 4
    assert lib.ffc_tail_workspace_bytes(0, 512, C.byref(need)) != 0
    a = _capi.TailArgs()
    assert lib.ffc_tail_forward(C.byref(a), None) != 0
    assert b'n_rows' in lib.ffc_last_error()
    a = _capi.TailArgs(0x1000, 0x1000, 64, 0x2000, 4, 64, 0, 1e-5, 0.1)
    assert lib.ffc_tail_forward(C.byref(a), None) != 0 and b'alias' in lib.ffc_last_error()
    a = _capi.TailArgs(0x1000, 0x3000, 32, 0x2000, 4, 64, 0, 1e-5, 0.1)
    assert lib.ffc_tail_forward(C.byref(a), None) != 0 and b'p_stride' in lib.ffc_last_error()
    a = _capi.TailArgs(0x1000, 0x3000, 64, 0x2000, 1, 64, 2, 1e-5, 0.1, None, None, None, None, 0x4000, 0x5000)
    assert lib.ffc_tail_forward(C.byref(a), None) != 0 and b'more than 1 row' in lib.ffc_last_error()
    a = _capi.TailArgs(0x1000, 0x3000, 64, 0x2000, 4, 64, 2, 1e-5, 0.1, None, None, None, None, 0x4000, 0x5000)
    assert lib.ffc_tail_forward(C.byref(a), None) != 0 and b'workspace' in lib.ffc_last_error()
    a = _capi.TailArgs(0x1000, 0x3000, 64, 0x2000, 4, 64, 0, 1e-5, 0.1)
    assert lib.ffc_tail_backward(C.byref(a), 0x6000, 64, 0x7000, 0x8000, None, None) != 0 and b'NORMALIZE' in lib.ffc_last_error()
    assert lib.ffc_launch_count() == launches            # nothing was launched


def test_tail_module_is_a_batchnorm1d():
    """host-side drop-in properties of FFCTail (no device needed): state_dict keys, parameter sharing, frozen weight, mode flags"""
    import torch
    import torch.nn as nn
    import ffc_b200
    bn = nn.BatchNorm1d(32, eps=1e-05)
    nn.init.constant_(bn.weight, 1.0)
    bn.weight.requires_grad = False                       # resnet_arcface.py:100-101
    bn.eval()
    t = ffc_b200.FFCTail.from_batchnorm(bn)
    assert isinstance(t, nn.BatchNorm1d) and not t.training and t.eps == bn.eps and t.momentum == bn.momentum
    assert t.weight is bn.weight and t.bias is bn.bias and t.running_mean is bn.running_mean and t.num_batches_tracked is bn.num_batches_tracked
    assert list(t.state_dict().keys()) == list(bn.state_dict().keys())
    assert [p.requires_grad for p in t.parameters()] == [False, True]

    class Net(nn.Module):
        def __init__(self):
            super().__init__()
            self.fc = nn.Linear(8, 32)
            self.features = nn.BatchNorm1d(32)

    net = Net()
    keys = list(net.state_dict().keys())
    n_params = len(list(net.parameters()))
    assert ffc_b200.fuse_tail(net) is net and isinstance(net.features, ffc_b200.FFCTail)
    assert list(net.state_dict().keys()) == keys and len(list(net.parameters())) == n_params
    assert ffc_b200.fuse_tail(net).features is net.features          # idempotent
    with pytest.raises(TypeError):
        ffc_b200.fuse_tail(net, 'fc')
    with pytest.raises(ValueError):
        net.features(torch.randn(4, 31))
    if not torch.cuda.is_available():
        with pytest.raises(ffc_b200.FFCError):
            net.features(torch.randn(4, 32))                          # no CPU fallback
        with pytest.raises(ffc_b200.FFCError):
            ffc_b200.l2_normalize(torch.randn(4, 32))


def test_head_create_rejects_unsupported_widths_up_front():
    """argument validation precedes any device work: widths the tensor-core sweep is not built for fail at create, with the reason"""
    import ctypes as C
    from ffc_b200 import _capi
    lib = _capi.lib()
    h = C.c_void_p()
    for D, prec, frag in ((192, 0, b'64, 128, 256 or 512'), (130, 1, b'multiple of 4'), (1024, 1, b'<= 512')):
        cfg = _capi.HeadConfig(64, 1024, 1024, 0, D, 0, 32.0, 0.4, 3, prec)
        assert lib.ffc_head_create(C.byref(cfg), C.byref(h)) == 1 and frag in lib.ffc_last_error(), (D, lib.ffc_last_error())
    cfg = _capi.HeadConfig(64, 1024, 1024, 0, 128, 0, 80.0, 0.4, 3, 0)          # logit range 2 * scale must fit the fp32 exponent range (<= 154): s <= 77
    assert lib.ffc_head_create(C.byref(cfg), C.byref(h)) == 1 and b'scale' in lib.ffc_last_error()
    cfg = _capi.HeadConfig(64, 1024, 1024, 0, 128, 0, 32.0, 0.4, 11, 0)          # ffc.py:48: hard_neg <= 10
    assert lib.ffc_head_create(C.byref(cfg), C.byref(h)) == 1 and b'topk' in lib.ffc_last_error()
