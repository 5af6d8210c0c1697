"""Loader for the UNMODIFIED reference (only works where /root/reference is mounted, i.e. the build
container).  TEST INFRASTRUCTURE ONLY - used by oracle/make_golden.py and by the CPU tests that pin the
oracle against the live reference.  Nothing on the GPU box may need this module to succeed: every caller
checks `reference_available()` first.
"""
from __future__ import annotations

import copy
import json
import os
import sys
import types

REFERENCE_ROOT = os.environ.get('VIPNERF_REFERENCE_ROOT', '/root/reference')


def reference_available() -> bool:
    return os.path.isfile(os.path.join(REFERENCE_ROOT, 'src', 'models', 'VipNeRF01.py'))


def _stub(name, **attrs):
    if name in sys.modules:
        return sys.modules[name]
    m = types.ModuleType(name)
    m.__dict__.update(attrs)
    sys.modules[name] = m
    return m


def import_reference():
    """Puts the reference's src/ on sys.path (stubbing the third-party modules the hot path never touches,
    SURVEY.md appendix B) and returns its `models.ModelFactory.get_model`."""
    if not reference_available():
        raise RuntimeError(f'reference not mounted at {REFERENCE_ROOT}')
    src = os.path.join(REFERENCE_ROOT, 'src')
    if src not in sys.path:
        sys.path.insert(0, src)
    for name in ('skimage', 'skimage.io', 'skimage.transform', 'deepdiff', 'matplotlib', 'matplotlib.pyplot',
                 'skvideo', 'skvideo.io'):
        try:
            __import__(name)
        except Exception:
            _stub(name)
    try:
        import simplejson  # noqa: F401
    except Exception:
        _stub('simplejson', load=json.load, dump=json.dump, loads=json.loads, dumps=json.dumps)
    from models.ModelFactory import get_model  # the reference's own factory
    return get_model


def reference_configs(ndc: bool) -> dict:
    """The reference's committed training configs: LLFF (NDC) train0012, DTU (world space) train0042."""
    run = 'train0012' if ndc else 'train0042'
    with open(os.path.join(REFERENCE_ROOT, 'runs', 'training', run, 'Configs.json')) as f:
        cfg = json.load(f)
    cfg['device'] = None
    return cfg


def build_reference_model(state_dict, ndc: bool, coarse_only: bool = False, white_bkgd: bool = False):
    get_model = import_reference()
    cfg = copy.deepcopy(reference_configs(ndc))
    cfg['model']['white_bkgd'] = white_bkgd
    if coarse_only:
        # SURVEY.md section 8a note 9: a config without fine_mlp renders only with retraw=True.  The reference
        # constructor still reads configs['model']['fine_mlp']['predict_visibility'] (:19), so coarse-only
        # is expressed as fine_mlp with zero extra samples at the model level; here we simply keep both MLPs
        # and compare the *_coarse keys.
        pass
    model = get_model(cfg, None)
    model.load_state_dict(state_dict)
    return model.eval()
