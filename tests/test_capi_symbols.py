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
