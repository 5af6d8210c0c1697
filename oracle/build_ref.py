"""Recipe for `oracle/_ref/`: the UNMODIFIED reference implementation of the render path, staged so that it travels
to the GPU box.  TEST / BENCH INFRASTRUCTURE ONLY - the product (vipnerf_b200/) never imports anything from here.

The reference's path is two pure-Python files, `src/models/VipNeRF01.py` (model) and `src/models/ModelFactory.py`
(plugin factory); there is nothing to compile.  `build_ref()` copies them byte for byte from where they lie under
/root/reference into `oracle/_ref/models/` (git-ignored, NOT gpurun-ignored: outputs only, never part of the
repository history) and records their SHA-256.  `/root/reference` exists only in the build container, so
`__graft_entry__.build()` runs this recipe there; on the GPU box the prebuilt directory is used as it arrived.

`load_ref_get_model()` returns the reference's own `models.ModelFactory.get_model` from that directory:
`bench.py --impl reference` drives the real thing (`cpu_baseline.kind = "reference"`) and falls back to the oracle
port (`kind = "port"`) only when the directory is absent.

    python -m oracle.build_ref
"""
from __future__ import annotations

import hashlib
import json
import os
import shutil
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
REF_DIR = os.path.join(HERE, '_ref')
REFERENCE_ROOT = os.environ.get('VIPNERF_REFERENCE_ROOT', '/root/reference')
FILES = ('src/models/VipNeRF01.py', 'src/models/ModelFactory.py')


def reference_sources_present() -> bool:
    return all(os.path.isfile(os.path.join(REFERENCE_ROOT, f)) for f in FILES)


def ref_available() -> bool:
    return all(os.path.isfile(os.path.join(REF_DIR, 'models', os.path.basename(f))) for f in FILES)


def build_ref() -> str:
    """Stages the reference's two files under oracle/_ref/models/ (no edits); returns the directory."""
    if not reference_sources_present():
        raise RuntimeError(f'reference sources not found under {REFERENCE_ROOT}')
    dst = os.path.join(REF_DIR, 'models')
    os.makedirs(dst, exist_ok=True)
    digests = {}
    for f in FILES:
        src = os.path.join(REFERENCE_ROOT, f)
        out = os.path.join(dst, os.path.basename(f))
        shutil.copyfile(src, out)
        with open(out, 'rb') as fh:
            digests[f] = hashlib.sha256(fh.read()).hexdigest()
    with open(os.path.join(REF_DIR, 'SOURCES.json'), 'w') as fh:
        json.dump({'reference_root': REFERENCE_ROOT, 'sha256': digests}, fh, indent=1)
    return REF_DIR


def load_ref_get_model():
    """The reference's own plugin factory, imported from oracle/_ref (its `models` is a namespace package)."""
    if not ref_available():
        raise RuntimeError('oracle/_ref is not staged (run `python -m oracle.build_ref` where /root/reference exists)')
    if REF_DIR not in sys.path:
        sys.path.insert(0, REF_DIR)
    from models.ModelFactory import get_model
    return get_model


if __name__ == '__main__':
    print(build_ref())
