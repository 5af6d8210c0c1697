"""Generates tests/golden/visibility_prior.npz by running the UNMODIFIED reference
VisibilityWeightsComputer.compute_weights (/root/reference/src/prior_generators/visibility/
VisibilityMask02_NeRF_LLFF.py:27-35) on seeded synthetic frame pairs with real LLFF camera matrices.
Run in the build container only:  python oracle/make_golden_prior.py        TEST INFRASTRUCTURE ONLY.
"""
from __future__ import annotations

import importlib.util
import os
import sys
import types

import numpy

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import ref_loader  # noqa: E402

REF = ref_loader.REFERENCE_ROOT


def load_reference_class():
    for name in ('pandas', 'simplejson', 'skimage', 'skimage.io', 'tqdm'):
        try:
            __import__(name)
        except Exception:   # noqa: BLE001
            m = types.ModuleType(name)
            if name == 'tqdm':
                m.tqdm = lambda x, *a, **k: x
            sys.modules[name] = m
    path = os.path.join(REF, 'src', 'prior_generators', 'visibility', 'VisibilityMask02_NeRF_LLFF.py')
    spec = importlib.util.spec_from_file_location('ref_visibility_mask', path)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod.VisibilityWeightsComputer


def smooth_image(rng, h, w):
    """A smooth random texture (so that neighbouring planes give graded errors) with some hard edges."""
    base = rng.uniform(0, 255, size=(h // 4 + 2, w // 4 + 2, 3))
    img = numpy.kron(base, numpy.ones((4, 4, 1)))[:h, :w]
    img[h // 3: h // 2, w // 4: w // 2] = rng.uniform(0, 255, size=3)
    return numpy.round(img).astype('uint8')


def main():
    cls = load_reference_class()
    poses = numpy.loadtxt(os.path.join(REF, 'data/databases/NeRF_LLFF/data/train_test_sets/set03/video_poses01/fern.csv'),
                          delimiter=',').reshape(-1, 4, 4)
    rng = numpy.random.default_rng(5)
    arrays = {}
    cases = [(36, 48, 12, 3, 4, 1.2, 60.0, 10), (40, 56, 16, 0, 0, 0.8, 20.0, 10), (30, 44, 64, 7, 90, 1.0, 6.2, 4)]
    for ci, (h, w, planes, i1, i2, dmin, dmax, temp) in enumerate(cases):
        k = numpy.array([[0.8 * w, 0, w / 2], [0, 0.8 * w, h / 2], [0, 0, 1.0]])
        k2 = k.copy()
        k2[0, 0] *= 1.03
        f1, f2 = smooth_image(rng, h, w), smooth_image(rng, h, w)
        if ci == 0:
            f2 = f1.copy()        # neighbouring cameras looking at the same texture: a mix of visible / occluded pixels
        if ci == 1:
            f2 = f1.copy()        # identical cameras and frames: coordinates land exactly on pixel centres
        comp = cls({'num_depth_planes': planes, 'temperature': temp})
        wts = comp.compute_weights(f1, f2, poses[i1], poses[i2], k, k2, dmin, dmax)
        arrays.update({f'c{ci}.frame1': f1, f'c{ci}.frame2': f2, f'c{ci}.extrinsic1': poses[i1], f'c{ci}.extrinsic2': poses[i2],
                       f'c{ci}.intrinsic1': k, f'c{ci}.intrinsic2': k2, f'c{ci}.params': numpy.array([dmin, dmax, planes, temp]),
                       f'c{ci}.weights': wts})
        print(ci, wts.shape, wts.dtype, float(wts.min()), float(wts.max()), float((wts > 0.5).mean()))
    path = os.path.join(ROOT, 'tests', 'golden', 'visibility_prior.npz')
    numpy.savez_compressed(path, **arrays)
    print(f'visibility_prior.npz: {os.path.getsize(path) / 1024:.0f} KiB')


if __name__ == '__main__':
    main()
