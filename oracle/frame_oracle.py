"""CPU oracle (numpy) of the steps either side of the render path: test-frame ray generation and output
post-processing.  TEST INFRASTRUCTURE ONLY - only tests/, __graft_entry__.smoke() and bench.py's CPU legs may import
it; the product path is vipnerf_b200/DataPreprocessorFused01.py + csrc/frame_kernels.cu.

Restates, citing /root/reference/src/data_preprocessors/DataPreprocessor01.py:
  preprocess_poses (test mode) :929-945, recenter_poses :948-950, convert_pose_to_standard_coordinates :952-958,
  change_coordinate_system :988-999, get_rays :335-352, get_ndc_rays :355-373, get_view_dirs :376-378,
  create_test_data :776-864, retrieve_inference_outputs :866-894, post_process_image :1074-1078,
  post_process_depth :1080-1083.
Pinned: tests/golden/frame_*.npz are outputs of the UNMODIFIED reference class (oracle/make_golden.py), and
tests/test_frame_oracle.py holds this file to them.
"""
from __future__ import annotations

from typing import Dict, List, Optional

import numpy


def preprocess_test_poses(poses: numpy.ndarray, translation_scale, average_pose: numpy.ndarray) -> numpy.ndarray:
    """:929-945 with train_mode=False (no bounds, no spherify)."""
    poses = numpy.array(poses, copy=True)
    poses[:, :3, 3] *= translation_scale                               # :931-932
    poses = average_pose[None] @ numpy.linalg.inv(poses)               # recenter_poses :948-950
    p = numpy.eye(3)                                                   # :954-956
    p[1, 1] = -1
    p[2, 2] = -1
    out = []
    for pose in poses:                                                 # change_coordinate_system :988-999
        rc = p.T @ pose[:3, :3] @ p
        tc = p @ pose[:3, 3:]
        out.append(numpy.concatenate([numpy.concatenate([rc, tc], axis=1), pose[3:]], axis=0))
    return numpy.stack(out).astype(numpy.float32)                      # :944


def get_rays(resolution, intrinsic: numpy.ndarray, pose: numpy.ndarray):
    """:335-352 (no mip-NeRF half-pixel shift)."""
    h, w = resolution
    x, y = numpy.meshgrid(numpy.arange(w, dtype=numpy.float32), numpy.arange(h, dtype=numpy.float32), indexing='xy')
    points_homo = numpy.stack([x, y, numpy.ones_like(x)], axis=2)
    dirs = (numpy.linalg.inv(intrinsic)[None, None] @ points_homo[:, :, :, None])[:, :, :, 0]
    dirs[:, :, 1:] *= -1
    rays_d = numpy.sum(dirs[..., numpy.newaxis, :] * pose[:3, :3], -1)
    rays_o = numpy.broadcast_to(pose[:3, -1], numpy.shape(rays_d))
    return rays_o, rays_d


def get_ndc_rays(rays_o, rays_d, resolution, intrinsic, near):
    """:355-373."""
    h, w = resolution
    fx, fy = intrinsic[0, 0], intrinsic[1, 1]
    t = -(near + rays_o[..., 2]) / rays_d[..., 2]
    rays_o = rays_o + t[..., None] * rays_d
    o0 = -1. / (w / (2. * fx)) * rays_o[..., 0] / rays_o[..., 2]
    o1 = -1. / (h / (2. * fy)) * rays_o[..., 1] / rays_o[..., 2]
    o2 = 1. + 2. * near / rays_o[..., 2]
    d0 = -1. / (w / (2. * fx)) * (rays_d[..., 0] / rays_d[..., 2] - rays_o[..., 0] / rays_o[..., 2])
    d1 = -1. / (h / (2. * fy)) * (rays_d[..., 1] / rays_d[..., 2] - rays_o[..., 1] / rays_o[..., 2])
    d2 = -2. * near / rays_o[..., 2]
    return numpy.stack([o0, o1, o2], -1), numpy.stack([d0, d1, d2], -1)


def get_view_dirs(rays_d):
    """:376-378."""
    return rays_d / numpy.linalg.norm(rays_d, ord=2, axis=-1, keepdims=True)


def create_test_data(model_configs: dict, ndc: bool, pose: numpy.ndarray, view_pose: Optional[numpy.ndarray] = None,
                     secondary_poses: Optional[List[numpy.ndarray]] = None, preprocess_pose: bool = True,
                     intrinsic: Optional[numpy.ndarray] = None, view_intrinsic: Optional[numpy.ndarray] = None
                     ) -> Dict[str, numpy.ndarray]:
    """:776-864 (without the torch wrapping / device move); arrays are flattened to [h*w, .] like the reference."""
    mc = model_configs
    avg = numpy.array(mc['average_pose'])
    if preprocess_pose:
        processed_pose = preprocess_test_poses(pose.copy()[None], mc['translation_scale'], avg)[0]
    else:
        processed_pose = pose.astype('float32')
    resolution = mc['resolution']
    intrinsic = (numpy.array(mc['intrinsic']) if intrinsic is None else intrinsic).astype('float32')
    rays_o, rays_d = get_rays(resolution, intrinsic, processed_pose)
    if view_pose is not None:
        processed_view_pose = preprocess_test_poses(view_pose.copy()[None], mc['translation_scale'], avg)[0]
        view_intrinsic = (numpy.array(mc['intrinsic']) if view_intrinsic is None else view_intrinsic).astype('float32')
        _, view_rays_d = get_rays(resolution, view_intrinsic, processed_view_pose)
        view_dirs = get_view_dirs(view_rays_d)
    else:
        view_dirs = get_view_dirs(rays_d)
    near = mc['near'] * numpy.ones_like(rays_d[..., :1])
    far = mc['far'] * numpy.ones_like(rays_d[..., :1])
    batch = {'rays_o': rays_o.copy().reshape(-1, 3), 'rays_d': rays_d.reshape(-1, 3),
             'view_dirs': view_dirs.reshape(-1, 3), 'near': near.reshape(-1, 1), 'far': far.reshape(-1, 1)}
    if ndc:
        o_ndc, d_ndc = get_ndc_rays(rays_o, rays_d, resolution, intrinsic, mc['near'])
        batch['rays_o_ndc'] = o_ndc.reshape(-1, 3)
        batch['rays_d_ndc'] = d_ndc.reshape(-1, 3)
        batch['near_ndc'] = (mc['near_ndc'] * numpy.ones_like(rays_d[..., :1])).reshape(-1, 1)
        batch['far_ndc'] = (mc['far_ndc'] * numpy.ones_like(rays_d[..., :1])).reshape(-1, 1)
    if secondary_poses is not None:
        sp = preprocess_test_poses(numpy.array([p.copy() for p in secondary_poses]), mc['translation_scale'], avg)
        sec_intrinsic = numpy.array(mc['intrinsic']).astype('float32')
        o2 = [numpy.array(get_rays(resolution, sec_intrinsic, s)[0]).reshape(-1, 3) for s in sp]
        batch['rays_o2'] = numpy.stack(o2, axis=1)
    return batch


def post_process_image(rgb):
    """:1074-1078."""
    return numpy.round(numpy.clip(rgb, a_min=0, a_max=1) * 255).astype('uint8')


def post_process_depth(depth):
    """:1080-1083."""
    return numpy.clip(depth, a_min=0, a_max=numpy.inf).astype('float32')


def retrieve_inference_outputs(outputs: Dict[str, numpy.ndarray], resolution, ndc: bool, suffix: str = '_fine'):
    """:866-894 on host arrays."""
    h, w = resolution
    ret = {'image': post_process_image(outputs[f'rgb{suffix}'].reshape(h, w, 3)),
           'depth': post_process_depth(outputs[f'depth{suffix}'].reshape(h, w)),
           'depth_var': post_process_depth(outputs[f'depth_var{suffix}'].reshape(h, w))}
    if ndc:
        ret['depth_ndc'] = post_process_depth(outputs[f'depth_ndc{suffix}'].reshape(h, w))
        ret['depth_var_ndc'] = post_process_depth(outputs[f'depth_var_ndc{suffix}'].reshape(h, w))
    if f'visibility2{suffix}' in outputs:
        v2 = outputs[f'visibility2{suffix}'].reshape((h, w, -1)).transpose([2, 0, 1])
        ret['visibility2'] = v2.astype('float32')
    return ret
