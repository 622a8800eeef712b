"""Drop-in for the reference ``lru.LRU`` (lru.py:21-255) backed by the device-resident cache.

Same surface (``get``, ``try_get``, ``view``, ``__contains__``, ``rollback_steps``, ``rollback_one_step``,
``state_dict``, ``restore``, ``keys``, ``__iter__``, ``clear``, ``capacity``, ``cur_idx``) and the same results,
slot for slot.  The scalar methods launch on the device and synchronise to return Python ints (API parity; the
FFC head uses the batched :meth:`assign` / :meth:`view_batch` which never synchronise).
"""
from __future__ import annotations

import ctypes as C

import torch

from . import _capi
from ._capi import check, ptr


class LRU:
    def __init__(self, capacity, device=None, journal_capacity=0):
        if not torch.cuda.is_available():
            raise _capi.FFCError('ffc_b200.LRU needs a CUDA device (there is no CPU fallback)')
        self.capacity = int(capacity)
        self.device = torch.device('cuda', torch.cuda.current_device()) if device is None else torch.device(device)
        self._lib = _capi.lib()
        h = C.c_void_p()
        with torch.cuda.device(self.device):
            check(self._lib.ffc_lru_create(self.capacity, int(journal_capacity), C.byref(h)))
        self._h = h
        self._key1 = torch.zeros(1, dtype=torch.int64, device=self.device)
        self._slot1 = torch.zeros(1, dtype=torch.int32, device=self.device)

    def __del__(self):
        h = self.__dict__.pop('_h', None)
        if h:
            try:
                self._lib.ffc_lru_destroy(h)
            except Exception:
                pass

    # -- batched, sync-free ---------------------------------------------------------------------
    def _s(self):
        return torch.cuda.current_stream(self.device).cuda_stream

    def assign(self, keys, journal=False, qpos=None, rows=None, cols=None, hit=None, ones_list=None, n_ones=None, cmask=None, n_dev=None):
        """B sequential get() (journal=False) / try_get() (journal=True) calls; see ffc_lru_assign."""
        n = keys.numel()
        assert keys.dtype == torch.int64 and keys.is_cuda and keys.is_contiguous()
        if cols is None:
            cols = torch.empty(n, dtype=torch.int32, device=self.device)
        s = self._s()
        for a in range(0, n, _capi.LRU_MAX_KEYS):       # one launch pair per call, whatever the batch size
            m = min(_capi.LRU_MAX_KEYS, n - a)
            off = lambda t, sz: None if t is None else t.data_ptr() + a * sz
            check(self._lib.ffc_lru_assign(self._h, keys.data_ptr() + a * 8, m, 1 if journal else 0, ptr(qpos), off(rows, 4), off(cols, 4),
                                           off(hit, 1), ptr(ones_list), ptr(n_ones), ptr(cmask), ptr(n_dev), a, s))
        return cols

    def view_batch(self, keys, out=None):
        n = keys.numel()
        assert keys.dtype == torch.int64 and keys.is_cuda and keys.is_contiguous()
        if out is None:
            out = torch.empty(n, dtype=torch.int32, device=self.device)
        check(self._lib.ffc_lru_view(self._h, keys.data_ptr(), n, out.data_ptr(), self._s()))
        return out

    def undo(self, steps, qpos=None):
        check(self._lib.ffc_lru_undo(self._h, int(steps), ptr(qpos), self._s()))

    # -- reference surface ------------------------------------------------------------------------
    def _check_key(self, key):
        key = int(key)
        if key in _capi.KEY_RESERVED or not (-(1 << 63) <= key < (1 << 63)):
            raise ValueError(f'key {key} is reserved / out of int64 range')
        return key

    def _one(self, key, journal):
        self._key1.fill_(self._check_key(key))
        self.assign(self._key1, journal=journal, cols=self._slot1)
        return int(self._slot1.item())

    def get(self, key):          # lru.py:44-89
        return self._one(key, False)

    def try_get(self, key):      # lru.py:157-204
        return self._one(key, True)

    def view(self, key):         # lru.py:147-151
        self._key1.fill_(self._check_key(key))
        return int(self.view_batch(self._key1, self._slot1).item())

    def __contains__(self, key):  # lru.py:145-146
        return self.view(key) >= 0

    def rollback_steps(self, steps):   # lru.py:252-255
        self.undo(steps)

    def rollback_one_step(self):       # lru.py:210-248
        self.undo(1)

    def _sizes(self):
        cur, jl = C.c_int64(), C.c_int64()
        check(self._lib.ffc_lru_size(self._h, C.byref(cur), C.byref(jl), self._s()))
        return cur.value, jl.value

    @property
    def cur_idx(self):
        return self._sizes()[0]

    @property
    def journal_len(self):
        return self._sizes()[1]

    def state_dict(self):        # lru.py:102-108: [(key, slot)] most -> least recently used
        keys = torch.empty(self.capacity, dtype=torch.int64)
        slots = torch.empty(self.capacity, dtype=torch.int32)
        n = C.c_int64()
        check(self._lib.ffc_lru_export(self._h, keys.data_ptr(), slots.data_ptr(), C.byref(n), self._s()))
        return list(zip(keys[:n.value].tolist(), slots[:n.value].tolist()))

    def restore(self, kvs):      # lru.py:113-128
        assert len(kvs) <= self.capacity
        keys = torch.tensor([self._check_key(k) for k, _ in kvs], dtype=torch.int64)
        slots = torch.tensor([int(v) for _, v in kvs], dtype=torch.int32)
        check(self._lib.ffc_lru_import(self._h, keys.data_ptr(), slots.data_ptr(), len(kvs), self._s()))

    def restore_arrays(self, keys, slots):
        """restore() from CPU int64 / int32 tensors (most- to least-recent), for large caches."""
        keys = keys.to(dtype=torch.int64, device='cpu').contiguous()
        slots = slots.to(dtype=torch.int32, device='cpu').contiguous()
        assert keys.numel() == slots.numel() <= self.capacity
        check(self._lib.ffc_lru_import(self._h, keys.data_ptr(), slots.data_ptr(), keys.numel(), self._s()))

    def __iter__(self):          # lru.py:94-98
        return iter(self.state_dict())

    def keys(self):              # lru.py:152-153
        return [k for k, _ in self.state_dict()]

    def clear(self):             # lru.py:132-141 (also resets cur_idx; the reference's clear() leaves it stale)
        check(self._lib.ffc_lru_clear(self._h, self._s()))
