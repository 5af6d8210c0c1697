"""Visibility-prior generator with the plane-sweep volume evaluated on the GPU.

Mirrors the class the reference's generator script is built around
(src/prior_generators/visibility/VisibilityMask02_NeRF_LLFF.py:22-39): `VisibilityWeightsComputer(configs)` with
`configs['num_depth_planes']` / `configs['temperature']`, and
`compute_weights(frame1, frame2, extrinsic1, extrinsic2, intrinsic1, intrinsic2, min_depth, max_depth) -> [h, w] float64`
plus the `weights > 0.5` mask rule of `start_generation` (:276-277).  The small matrix algebra (inverse intrinsics, the
relative transformation, the inverse-depth planes) uses the reference's numpy expressions on the host; the per-pixel
plane sweep runs in `vipnerf_visibility_prior`.  No CPU fallback.
"""
from __future__ import annotations

import ctypes
from typing import Optional

import numpy
import torch

from . import _lib


class VisibilityWeightsComputer:
    def __init__(self, configs: dict, device: Optional[torch.device] = None):
        self.configs = configs
        if not torch.cuda.is_available():
            raise RuntimeError('VisibilityPriorFused needs a CUDA device; there is no CPU fallback')
        self.device = torch.device('cuda', torch.cuda.current_device()) if device is None else torch.device(device)

    @staticmethod
    def get_depth_planes(min_depth, max_depth, num_depth_planes):
        return 1 / numpy.linspace(1 / min_depth, 1 / max_depth, num_depth_planes)    # :37-39

    def compute_weights_device(self, frame1, frame2, extrinsic1, extrinsic2, intrinsic1, intrinsic2, min_depth,
                               max_depth, with_mask: bool = True):
        """Device tensors (weights fp64 [h, w], mask bool [h, w] or None); frames may be numpy uint8 arrays or CUDA
        uint8 tensors."""
        lib = _lib.load()

        def to_dev(f):
            t = torch.from_numpy(numpy.ascontiguousarray(f)) if isinstance(f, numpy.ndarray) else f
            if t.dtype != torch.uint8 or t.dim() != 3 or t.shape[2] != 3:
                raise ValueError('frames must be uint8 [h, w, 3]')
            return t.to(self.device).contiguous()

        f1, f2 = to_dev(frame1), to_dev(frame2)
        if f1.shape != f2.shape:
            raise ValueError(f'frame shapes differ: {tuple(f1.shape)} vs {tuple(f2.shape)}')
        h, w = int(f1.shape[0]), int(f1.shape[1])
        if intrinsic2 is None:
            intrinsic2 = numpy.copy(intrinsic1)                                       # :57-58
        planes = self.get_depth_planes(min_depth, max_depth, self.configs['num_depth_planes']).astype(numpy.float64)
        t = numpy.matmul(extrinsic2, numpy.linalg.inv(extrinsic1)).astype(numpy.float64)            # :59
        k1inv = numpy.linalg.inv(intrinsic1).astype(numpy.float64)                                   # :67
        k2 = numpy.asarray(intrinsic2, dtype=numpy.float64)
        weights = torch.empty((h, w), dtype=torch.float64, device=self.device)
        mask = torch.empty((h, w), dtype=torch.uint8, device=self.device) if with_mask else None

        def dptr(a):
            a = numpy.ascontiguousarray(a, dtype=numpy.float64)
            return a, a.ctypes.data_as(ctypes.POINTER(ctypes.c_double))

        keep = [dptr(k1inv), dptr(t), dptr(k2), dptr(planes)]
        with torch.cuda.device(self.device):
            stream = torch.cuda.current_stream(self.device).cuda_stream
            _lib.check(lib.vipnerf_visibility_prior(h, w, f1.data_ptr(), f2.data_ptr(), keep[0][1], keep[1][1], keep[2][1],
                                                    keep[3][1], len(planes), float(self.configs['temperature']),
                                                    weights.data_ptr(), mask.data_ptr() if with_mask else None, stream),
                       'vipnerf_visibility_prior')
        return weights, (mask.bool() if with_mask else None)

    def compute_weights(self, frame1: numpy.ndarray, frame2: numpy.ndarray, extrinsic1, extrinsic2, intrinsic1,
                        intrinsic2, min_depth: float, max_depth: float) -> numpy.ndarray:
        """The reference's signature and return type (:27-35)."""
        weights, _ = self.compute_weights_device(frame1, frame2, extrinsic1, extrinsic2, intrinsic1, intrinsic2,
                                                 min_depth, max_depth, with_mask=False)
        return weights.cpu().numpy()
