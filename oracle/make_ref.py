#!/usr/bin/env python
"""Recipe for ``oracle/_ref/``: the UNMODIFIED reference head, staged for the GPU box.

TEST INFRASTRUCTURE ONLY.  `/root/reference` exists in the build container but not on the GPU box, and the
reference is a plain script tree (pure Python, no build step), so "compiling the reference
where it lies" degenerates to staging the files its hot path imports:

    lru.py  ffc.py  model/{__init__,mobilefacenet_def,resnet_arcface,resnet_std}.py

byte for byte into ``oracle/_ref/`` -- a directory that is git-ignored (never part of the history: no reference source
is committed) but NOT gpurun-ignored, so it travels to the GPU box like the built ``.so``.  ``__graft_entry__.build()``
runs this whenever `/root/reference` is present.  Consumers: ``oracle/ref_shim.py`` (falls back to ``oracle/_ref`` when
`/root/reference` is absent), i.e. ``bench.py``'s ``cpu_baseline`` / ``--impl reference`` legs (``kind: "reference"``) and
the ``-m gpu`` tests that drive the reference's own backbones (``model/``) under ``ffc_b200.FFC``.  The product package
never reads it.  A MANIFEST with the SHA-256 of every staged file is written next to them.
"""
from __future__ import annotations

import hashlib
import json
import os
import shutil
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
DST = os.path.join(HERE, '_ref')
FILES = ['lru.py', 'ffc.py', 'model/__init__.py', 'model/mobilefacenet_def.py', 'model/resnet_arcface.py', 'model/resnet_std.py']


def stage(src_root: str = '/root/reference') -> str | None:
    if not os.path.isfile(os.path.join(src_root, 'ffc.py')):
        return None
    manifest = {}
    for rel in FILES:
        src, dst = os.path.join(src_root, rel), os.path.join(DST, rel)
        os.makedirs(os.path.dirname(dst), exist_ok=True)
        shutil.copyfile(src, dst)
        os.chmod(dst, 0o644)
        manifest[rel] = hashlib.sha256(open(dst, 'rb').read()).hexdigest()
    with open(os.path.join(DST, 'MANIFEST.json'), 'w') as f:
        json.dump(dict(source=src_root, files=manifest), f, indent=1, sort_keys=True)
    return DST


if __name__ == '__main__':
    out = stage(sys.argv[1] if len(sys.argv) > 1 else '/root/reference')
    print(out if out else 'reference tree not present: nothing staged')
