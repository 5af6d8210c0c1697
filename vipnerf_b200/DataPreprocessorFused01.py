"""Test-mode data preprocessor with the per-pixel work on the GPU: the steps either side of the render path
(SURVEY.md section 8, row f3).

Plugin contract (reference src/data_preprocessors/DataPreprocessorFactory.py:13-26): file `<Name>NN.py` holding
class `<Name>(configs, mode, raw_data_dict, model_configs)`; here
`configs['data_loader']['data_preprocessor_name'] = 'DataPreprocessorFused01'`.  It mirrors the two methods the
reference's Tester calls (src/Tester01.py:57-66, :203-211):

  create_test_data(pose, view_pose, secondary_poses, preprocess_pose, intrinsic, view_intrinsic, secondary_intrinsics)
      -> dict of CUDA tensors with the keys of DataPreprocessor.create_test_data
         (src/data_preprocessors/DataPreprocessor01.py:776-864).  The 4x4 pose algebra (preprocess_poses :906-945,
         recenter_poses :948-950, convert_pose_to_standard_coordinates :952-958) runs on the host as in the reference
         - it is a handful of 4x4 products; get_rays / get_view_dirs / get_ndc_rays and the near/far/rays_o2 fills
         run in `vipnerf_generate_rays` directly into device memory, so no per-ray data crosses PCIe.
  retrieve_inference_outputs(network_outputs) -> dict of numpy arrays (image uint8 [h,w,3], depth maps, visibility2
         [V,h,w]) like :866-894, with clip / round / uint8 / transpose done by `vipnerf_postprocess_frame` and only
         the finished frame copied to the host.

Training / validation modes (ray caches, sparse-depth and visibility-prior batches) are not part of this build and
raise NotImplementedError; there is no CPU fallback.
"""
from __future__ import annotations

import ctypes
from typing import Dict, List, Optional

import numpy
import torch

from . import _lib


def preprocess_test_poses(poses: numpy.ndarray, translation_scale, average_pose: numpy.ndarray) -> numpy.ndarray:
    """preprocess_poses(train_mode=False) of the reference (DataPreprocessor01.py:929-945) for configs without
    `spherify`: scale the translations, re-centre on the average pose (`avg @ inv(pose)`, :948-950) and convert the
    (x, -y, -z) convention to NeRF's (:952-958, :988-999).  Same numpy operations in the same order and dtypes."""
    poses = numpy.array(poses, copy=True)
    poses[:, :3, 3] *= translation_scale
    poses = average_pose[None] @ numpy.linalg.inv(poses)
    perm = numpy.eye(3)
    perm[1, 1] = -1
    perm[2, 2] = -1
    changed = []
    for pose in poses:
        rc = perm.T @ pose[:3, :3] @ perm
        tc = perm @ pose[:3, 3:]
        changed.append(numpy.concatenate([numpy.concatenate([rc, tc], axis=1), pose[3:]], axis=0))
    return numpy.stack(changed).astype(numpy.float32)


class DataPreprocessorFused:
    def __init__(self, configs: dict, mode: str, raw_data_dict: Optional[dict] = None,
                 model_configs: Optional[dict] = None):
        self.configs = configs
        self.mode = mode.lower()
        if self.mode != 'test':
            raise NotImplementedError('DataPreprocessorFused covers test-mode frame rendering only; use the reference '
                                      'DataPreprocessor01 for training / validation batches')
        if 'mip_nerf' in configs['data_loader']:
            raise NotImplementedError('mip-NeRF radii are not part of this build')
        if configs['data_loader'].get('spherify', False):
            raise NotImplementedError('spherify is not part of this build')
        self.ndc = configs['data_loader']['ndc']
        self.model_configs = model_configs
        device = configs.get('device')
        if isinstance(device, (list, tuple)):
            device = device[0] if len(device) > 0 else None
        if device is None or device == 'cpu':
            device = torch.cuda.current_device() if torch.cuda.is_available() else None
        if device is None:
            raise RuntimeError('DataPreprocessorFused needs a CUDA device; there is no CPU fallback')
        self.device = torch.device('cuda', int(device)) if not isinstance(device, torch.device) else device

    def get_model_configs(self):
        return self.model_configs

    # ------------------------------------------------------------------ create_test_data
    def _processed(self, poses: List[numpy.ndarray], preprocess_pose: bool) -> numpy.ndarray:
        if not preprocess_pose:
            return numpy.stack([p.astype('float32') for p in poses])
        return preprocess_test_poses(numpy.stack(poses), self.model_configs['translation_scale'],
                                     numpy.array(self.model_configs['average_pose']))

    def camera(self, pose, view_pose=None, secondary_poses=None, preprocess_pose=True, intrinsic=None,
               view_intrinsic=None) -> _lib.Camera:
        mc = self.model_configs
        h, w = mc['resolution']
        cam = _lib.Camera()
        cam.height, cam.width, cam.ndc = int(h), int(w), int(bool(self.ndc))
        intrinsic = (numpy.array(mc['intrinsic']) if intrinsic is None else intrinsic).astype('float32')
        pose_p = self._processed([pose.copy()], preprocess_pose)[0]
        cam.kinv[:] = numpy.linalg.inv(intrinsic).astype(numpy.float32).reshape(-1).tolist()
        cam.pose[:] = pose_p[:3, :4].reshape(-1).tolist()
        if view_pose is not None:
            view_intrinsic = (numpy.array(mc['intrinsic']) if view_intrinsic is None else view_intrinsic).astype('float32')
            # the reference always pre-processes the view pose (:801-807)
            vp = preprocess_test_poses(view_pose.copy()[None], mc['translation_scale'], numpy.array(mc['average_pose']))[0]
            cam.has_view_pose = 1
            cam.view_kinv[:] = numpy.linalg.inv(view_intrinsic).astype(numpy.float32).reshape(-1).tolist()
            cam.view_pose[:] = vp[:3, :4].reshape(-1).tolist()
        cam.near, cam.far = float(mc['near']), float(mc['far'])
        if self.ndc:
            cam.near_ndc, cam.far_ndc = float(mc['near_ndc']), float(mc['far_ndc'])
            fx, fy = intrinsic[0, 0], intrinsic[1, 1]
            cam.sx = float(numpy.float32(-1. / (w / (2. * fx))))   # :364-365, evaluated with the reference's dtypes
            cam.sy = float(numpy.float32(-1. / (h / (2. * fy))))
        if secondary_poses is not None:
            # secondary poses are always pre-processed (:842-847); only their camera centre is used (:851)
            sp = preprocess_test_poses(numpy.array([p.copy() for p in secondary_poses]), mc['translation_scale'],
                                       numpy.array(mc['average_pose']))
            if len(sp) > 8:
                raise NotImplementedError('at most 8 secondary views')
            cam.n_sec_views = len(sp)
            cam.sec_origins[:3 * len(sp)] = sp[:, :3, 3].reshape(-1).tolist()
        return cam

    def create_test_data(self, pose: numpy.ndarray, view_pose: Optional[numpy.ndarray] = None,
                         secondary_poses: Optional[List[numpy.ndarray]] = None, preprocess_pose: bool = True,
                         intrinsic: Optional[numpy.ndarray] = None, view_intrinsic: Optional[numpy.ndarray] = None,
                         secondary_intrinsics: Optional[List[numpy.ndarray]] = None,
                         first_pixel: int = 0, n_rays: Optional[int] = None) -> Dict[str, torch.Tensor]:
        """Ray batch of a whole frame (or of pixels [first_pixel, first_pixel + n_rays) - used to shard a frame over
        GPUs).  `secondary_intrinsics` is accepted for signature compatibility; like in the reference only the
        secondary cameras' centres enter the batch."""
        cam = self.camera(pose, view_pose, secondary_poses, preprocess_pose, intrinsic, view_intrinsic)
        return self.generate(cam, first_pixel, n_rays)

    def generate(self, cam: _lib.Camera, first_pixel: int = 0, n_rays: Optional[int] = None) -> Dict[str, torch.Tensor]:
        lib = _lib.load()
        R = cam.height * cam.width - first_pixel if n_rays is None else int(n_rays)
        V = cam.n_sec_views
        widths = {'rays_o': 3, 'rays_d': 3, 'view_dirs': 3, 'near': 1, 'far': 1}
        if cam.ndc:
            widths.update({'rays_o_ndc': 3, 'rays_d_ndc': 3, 'near_ndc': 1, 'far_ndc': 1})
        # one allocation, one launch: every key is a contiguous slice of the same buffer
        # (every slice starts on a 16-byte boundary, which the render entry points require)
        if V > 0:
            widths['rays_o2'] = 3 * V
        pad4 = lambda n: (n + 3) // 4 * 4
        flat = torch.empty(max(sum(pad4(R * wd) for wd in widths.values()), 4), dtype=torch.float32, device=self.device)
        batch, bufs, off = {}, _lib.RayBuffers(), 0
        for name, wd in widths.items():
            t = flat[off:off + R * wd]
            batch[name] = t.view(R, V, 3) if name == 'rays_o2' else t.view(R, wd)
            setattr(bufs, name, batch[name].data_ptr())
            off += pad4(R * wd)
        with torch.cuda.device(self.device):
            stream = torch.cuda.current_stream(self.device).cuda_stream
            _lib.check(lib.vipnerf_generate_rays(ctypes.byref(cam), first_pixel, R, ctypes.byref(bufs), stream),
                       'vipnerf_generate_rays')
        return batch

    # ------------------------------------------------------------------ retrieve_inference_outputs
    def retrieve_inference_outputs(self, network_outputs: dict):
        h, w = self.model_configs['resolution']
        if 'fine_mlp' in self.configs['model']:
            suffix = '_fine'
        elif 'coarse_mlp' in self.configs['model']:
            suffix = '_coarse'
        else:
            raise RuntimeError
        frame = self.postprocess(network_outputs, suffix)
        out = {'image': frame['image'].reshape(h, w, 3)}
        for key in ('depth', 'depth_var', 'depth_ndc', 'depth_var_ndc'):
            if key in frame:
                out[key] = frame[key].reshape(h, w)
        if 'visibility2' in frame:
            out['visibility2'] = frame['visibility2'].reshape(-1, h, w)
        return out

    def postprocess(self, network_outputs: dict, suffix: str) -> Dict[str, numpy.ndarray]:
        """Device part of retrieve_inference_outputs for any number of rays; returns flat host arrays."""
        lib = _lib.load()
        rgb = network_outputs[f'rgb{suffix}']
        if not rgb.is_cuda:
            raise RuntimeError('DataPreprocessorFused needs the CUDA outputs of the model; there is no CPU fallback')
        R = rgb.shape[0]
        depth_keys = ['depth', 'depth_var'] + (['depth_ndc', 'depth_var_ndc'] if self.ndc else [])
        depth_in = [network_outputs[f'{k}{suffix}'].contiguous() for k in depth_keys]
        vis2 = network_outputs.get(f'visibility2{suffix}')
        V = vis2.shape[1] if vis2 is not None else 0
        # one device buffer for the whole finished frame: [depth maps | visibility2^T | image bytes]
        n_f32 = R * (len(depth_keys) + V)
        frame = torch.empty(n_f32 * 4 + R * 3, dtype=torch.uint8, device=rgb.device)
        f32 = frame[:n_f32 * 4].view(torch.float32)
        image = frame[n_f32 * 4:]
        depth_out = [f32[i * R:(i + 1) * R] for i in range(len(depth_keys))]
        vis2_out = f32[len(depth_keys) * R:] if V else None
        arr_in = (ctypes.c_void_p * len(depth_keys))(*[t.data_ptr() for t in depth_in])
        arr_out = (ctypes.c_void_p * len(depth_keys))(*[t.data_ptr() for t in depth_out])
        rgb_c = rgb.contiguous()
        vis2_c = vis2.contiguous() if vis2 is not None else None
        with torch.cuda.device(rgb.device):
            stream = torch.cuda.current_stream(rgb.device).cuda_stream
            _lib.check(lib.vipnerf_postprocess_frame(R, V, rgb_c.data_ptr(), image.data_ptr(), len(depth_keys), arr_in,
                                                     arr_out, vis2_c.data_ptr() if V else None,
                                                     vis2_out.data_ptr() if V else None, stream),
                       'vipnerf_postprocess_frame')
        host = frame.cpu().numpy()     # the single device-to-host copy of the frame
        host_f32 = host[:n_f32 * 4].view(numpy.float32)
        out = {'image': host[n_f32 * 4:].reshape(R, 3)}
        for i, k in enumerate(depth_keys):
            out[k] = host_f32[i * R:(i + 1) * R]
        if V:
            out['visibility2'] = host_f32[len(depth_keys) * R:].reshape(V, R)
        return out
