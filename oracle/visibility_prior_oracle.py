"""CPU oracle (numpy) of the visibility-prior generator: plane-sweep-volume visibility weights of frame 1 w.r.t.
frame 2.  TEST INFRASTRUCTURE ONLY (tests/, smoke, bench CPU legs); the product path is
vipnerf_b200/VisibilityPriorFused02.py + csrc/prior_kernels.cu.

Restates /root/reference/src/prior_generators/visibility/VisibilityMask02_NeRF_LLFF.py:
  compute_weights :27-35, get_depth_planes :37-39, create_psv :41-47, compute_transformed_coordinates :49-82,
  bilinear_interpolation :84-162 (with mask2 = flow12_mask = all ones, is_image=False), create_grid :164-171, and the
  mask rule `weights > 0.5` of start_generation :276-277.
Everything is float64 like the reference (its extrinsics / intrinsics are float64 and promote the whole pipeline).
Pinned: tests/golden/visibility_prior.npz holds outputs of the UNMODIFIED reference class
(oracle/make_golden_prior.py); tests/test_prior_oracle.py holds this file to them.
"""
from __future__ import annotations

import numpy


def get_depth_planes(min_depth, max_depth, num_depth_planes):
    """:37-39 - planes uniform in inverse depth."""
    return 1 / numpy.linspace(1 / min_depth, 1 / max_depth, num_depth_planes)


def transformed_coordinates(h, w, depth_planes, extrinsic1, extrinsic2, intrinsic1, intrinsic2):
    """:49-82 - where pixel (x, y) of camera 1 at each plane depth lands in camera 2; [h, w, d, 2] (x, y)."""
    transformation = numpy.matmul(extrinsic2, numpy.linalg.inv(extrinsic1))
    x2d, y2d = numpy.meshgrid(numpy.array(range(w)), numpy.array(range(h)))
    pos = numpy.stack([x2d, y2d, numpy.ones((h, w))], axis=2)[:, :, None, :, None]          # (h, w, 1, 3, 1)
    unnormalized = numpy.matmul(numpy.linalg.inv(intrinsic1)[None, None, None], pos)             # (h, w, 1, 3, 1)
    world = depth_planes[None, None, :, None, None] * unnormalized                             # (h, w, d, 3, 1)
    world_homo = numpy.concatenate([world, numpy.ones((h, w, len(depth_planes), 1, 1))], axis=3)
    trans = numpy.matmul(transformation[None, None, None], world_homo)[:, :, :, :3]             # (h, w, d, 3, 1)
    norm = numpy.matmul(intrinsic2[None, None, None], trans)
    return norm[:, :, :, :2, 0] / norm[:, :, :, 2:3, 0]


def bilinear_sample(frame2, trans_pos):
    """:84-162 with all-ones masks: frame2 [h, w, c] (zero-padded by one pixel), trans_pos [h, w, d, 2] -> [h, w, d, c].
    Note the reference's weights: at an exactly integer coordinate floor == ceil and all four weights are 1."""
    h, w, _ = frame2.shape
    off = trans_pos + 1
    fl = numpy.floor(off).astype('int')
    ce = numpy.ceil(off).astype('int')
    off = off.copy()
    for a, hi in ((0, w + 1), (1, h + 1)):
        off[..., a] = numpy.clip(off[..., a], a_min=0, a_max=hi)
        fl[..., a] = numpy.clip(fl[..., a], a_min=0, a_max=hi)
        ce[..., a] = numpy.clip(ce[..., a], a_min=0, a_max=hi)
    w_nw = (1 - (off[..., 1] - fl[..., 1])) * (1 - (off[..., 0] - fl[..., 0]))
    w_sw = (1 - (ce[..., 1] - off[..., 1])) * (1 - (off[..., 0] - fl[..., 0]))
    w_ne = (1 - (off[..., 1] - fl[..., 1])) * (1 - (ce[..., 0] - off[..., 0]))
    w_se = (1 - (ce[..., 1] - off[..., 1])) * (1 - (ce[..., 0] - off[..., 0]))
    f2 = numpy.pad(frame2, pad_width=((1, 1), (1, 1), (0, 0)), mode='constant', constant_values=0)
    m2 = numpy.pad(numpy.ones((h, w), dtype=bool), pad_width=((1, 1), (1, 1)), mode='constant', constant_values=0)
    nr, dr = 0, 0
    for wt, yy, xx in ((w_nw, fl[..., 1], fl[..., 0]), (w_sw, ce[..., 1], fl[..., 0]),
                       (w_ne, fl[..., 1], ce[..., 0]), (w_se, ce[..., 1], ce[..., 0])):
        m = m2[yy, xx][..., None]
        nr = nr + wt[..., None] * f2[yy, xx] * m
        dr = dr + wt[..., None] * m
    with numpy.errstate(divide='ignore', invalid='ignore'):
        return numpy.where(dr > 0, nr / dr, 0)


def compute_weights(frame1, frame2, extrinsic1, extrinsic2, intrinsic1, intrinsic2, min_depth, max_depth,
                    num_depth_planes=64, temperature=10):
    """:27-35 - exp(-min over planes of the mean absolute colour error / temperature), [h, w] float64."""
    h, w = frame1.shape[:2]
    planes = get_depth_planes(min_depth, max_depth, num_depth_planes)
    coords = transformed_coordinates(h, w, planes, extrinsic1, extrinsic2, intrinsic1, intrinsic2)
    grid = numpy.stack(numpy.meshgrid(numpy.arange(0, w), numpy.arange(0, h)), axis=2)[:, :, None, :]   # create_grid
    trans_pos = (coords - grid) + grid                     # create_psv forms the flow and bilinear_interpolation adds the grid back
    psv = bilinear_sample(frame2.astype('float32'), trans_pos)
    abs_error = numpy.mean(numpy.abs(psv - frame1[:, :, None, :]), axis=3)
    return numpy.exp(-numpy.min(abs_error, axis=2) / temperature)


def visibility_mask(weights):
    """start_generation :276-277."""
    return weights > 0.5
