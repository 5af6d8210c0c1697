"""ctypes binding of libvipnerf_b200.so (C ABI declared in include/vipnerf.h).

The library is the product: there is no Python / CPU fallback.  If the shared object is missing this module
raises immediately with the command that builds it.
"""
from __future__ import annotations

import ctypes
import os
import threading
from ctypes import POINTER, Structure, c_char_p, c_double, c_float, c_int, c_int32, c_int64, c_size_t, c_uint32, c_void_p

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, 'libvipnerf_b200.so')

ABI_VERSION = 1
FLAG_NDC, FLAG_WHITE_BKGD, FLAG_LINDISP, FLAG_TRAIN_TF32, FLAG_TRAIN_F16 = 1, 2, 4, 8, 16
TRAIN_PRECISIONS = ('fp32', 'tf32', 'fp16')
PRECISION = {'fp32': 0, 'bf16': 1, 'bf16x3': 2, 'fp16': 3}
STATUS_NAMES = {0: 'OK', -1: 'EINVAL', -2: 'EUNSUPPORTED', -3: 'ECUDA', -4: 'EWORKSPACE', -5: 'EABI'}

c_float_p = c_void_p  # device pointers travel as integers


class Cfg(Structure):
    _fields_ = [('abi', c_int32), ('n_coarse', c_int32), ('n_fine', c_int32), ('l_pts', c_int32),
                ('l_view', c_int32), ('depth', c_int32), ('width', c_int32), ('skip', c_int32),
                ('n_sec_views', c_int32), ('flags', c_uint32), ('precision', c_int32), ('reserved', c_int32)]


RAY_FIELDS = ('rays_o', 'rays_d', 'view_dirs', 'near', 'far', 'rays_o_ndc', 'rays_d_ndc', 'near_ndc', 'far_ndc',
              'rays_o2', 't_vals', 'u_vals', 't_rand', 'u_rand')
PASS_FIELDS = ('rgb', 'acc', 'depth', 'depth_var', 'depth_ndc', 'depth_var_ndc', 'visibility2', 'alpha', 'z_vals',
               'visibility', 'weights', 'raw_sigma', 'raw_rgb', 'raw_visibility', 'raw_visibility2')


class Rays(Structure):
    _fields_ = [(name, c_float_p) for name in RAY_FIELDS]


class PassOut(Structure):
    _fields_ = [(name, c_float_p) for name in PASS_FIELDS]


class Out(Structure):
    _fields_ = [('coarse', PassOut), ('fine', PassOut)]


class Camera(Structure):   # vipnerf_camera
    _fields_ = [('height', c_int32), ('width', c_int32), ('ndc', c_int32), ('n_sec_views', c_int32),
                ('has_view_pose', c_int32), ('kinv', c_float * 9), ('pose', c_float * 12), ('view_kinv', c_float * 9),
                ('view_pose', c_float * 12), ('near', c_float), ('far', c_float), ('near_ndc', c_float),
                ('far_ndc', c_float), ('sx', c_float), ('sy', c_float), ('sec_origins', c_float * 24)]


class LossSpec(Structure):   # vipnerf_loss_spec
    _fields_ = [('target_rgb', c_void_p), ('mask_nerf', c_void_p), ('mask_sparse_depth', c_void_p),
                ('sparse_depth', c_void_p), ('prior', c_void_p), ('losses_dev', c_void_p), ('w_mse', c_float),
                ('w_visibility', c_float), ('w_prior', c_float), ('w_sparse_depth', c_float)]


class GatherColumn(Structure):   # vipnerf_gather_column
    _fields_ = [('table', c_void_p), ('out', c_void_p), ('width', c_int32), ('row_classes', c_int32),
                ('fill_is_int', c_int32), ('reserved', c_int32)]


RAY_BUFFER_FIELDS = RAY_FIELDS[:10]


class RayBuffers(Structure):   # vipnerf_ray_buffers
    _fields_ = [(name, c_float_p) for name in RAY_BUFFER_FIELDS]


EXPORTS = {
    # name: (restype, argtypes)  -- must list every symbol include/vipnerf.h declares
    'vipnerf_abi_version': (c_int, []),
    'vipnerf_last_error': (c_char_p, []),
    'vipnerf_check_config': (c_int, [POINTER(Cfg)]),
    'vipnerf_packed_weight_bytes': (c_size_t, [POINTER(Cfg)]),
    'vipnerf_pack_weights': (c_int, [POINTER(Cfg), POINTER(c_void_p), c_void_p, c_void_p]),
    'vipnerf_workspace_bytes': (c_size_t, [POINTER(Cfg), c_int64]),
    'vipnerf_render_forward': (c_int, [POINTER(Cfg), POINTER(Rays), c_int64, c_void_p, c_void_p, POINTER(Out),
                                       c_void_p, c_size_t, c_void_p]),
    'vipnerf_mlp_forward': (c_int, [POINTER(Cfg), POINTER(Rays), c_int64, c_int32, c_void_p, c_void_p,
                                    POINTER(PassOut), c_void_p, c_size_t, c_void_p]),
    'vipnerf_coarse_z': (c_int, [POINTER(Cfg), POINTER(Rays), c_int64, c_void_p, c_void_p]),
    'vipnerf_composite': (c_int, [POINTER(Cfg), POINTER(Rays), c_int64, c_int32, c_void_p, c_void_p, c_void_p,
                                  c_void_p, POINTER(PassOut), c_void_p, c_void_p]),
    'vipnerf_generate_rays': (c_int, [POINTER(Camera), c_int64, c_int64, POINTER(RayBuffers), c_void_p]),
    'vipnerf_gather_train_batch': (c_int, [c_void_p, c_void_p, c_int64, POINTER(GatherColumn), c_int32, c_void_p]),
    'vipnerf_postprocess_frame': (c_int, [c_int64, c_int32, c_void_p, c_void_p, c_int32, POINTER(c_void_p),
                                          POINTER(c_void_p), c_void_p, c_void_p, c_void_p]),
    'vipnerf_visibility_prior': (c_int, [c_int32, c_int32, c_void_p, c_void_p, POINTER(c_double), POINTER(c_double),
                                         POINTER(c_double), POINTER(c_double), c_int32, c_double, c_void_p, c_void_p,
                                         c_void_p]),
    'vipnerf_train_saved_bytes': (c_size_t, [POINTER(Cfg), c_int64]),
    'vipnerf_train_workspace_bytes': (c_size_t, [POINTER(Cfg), c_int64]),
    'vipnerf_train_forward': (c_int, [POINTER(Cfg), POINTER(Rays), c_int64, c_void_p, c_void_p, c_void_p, c_void_p,
                                      POINTER(Out), c_void_p, c_size_t, c_void_p, c_size_t, c_void_p]),
    'vipnerf_train_backward': (c_int, [POINTER(Cfg), POINTER(Rays), c_int64, c_void_p, c_void_p, POINTER(Out),
                                       POINTER(Out), c_void_p, c_size_t, POINTER(c_void_p), POINTER(c_void_p),
                                       c_void_p, c_size_t, c_void_p]),
    'vipnerf_fused_losses': (c_int, [POINTER(Cfg), c_int64, POINTER(Out), POINTER(LossSpec), c_void_p, c_void_p, c_size_t,
                                     c_void_p]),
    'vipnerf_train_backward_fused': (c_int, [POINTER(Cfg), POINTER(Rays), c_int64, c_void_p, c_void_p, POINTER(Out),
                                             POINTER(Out), POINTER(LossSpec), c_void_p, c_void_p, c_size_t,
                                             POINTER(c_void_p), POINTER(c_void_p), c_void_p, c_size_t, c_void_p]),
    'vipnerf_composite_backward': (c_int, [POINTER(Cfg), POINTER(Rays), c_int64, c_int32, c_void_p, c_void_p, c_void_p,
                                           c_void_p, c_void_p, POINTER(PassOut), c_void_p, c_void_p, c_void_p]),
    'vipnerf_param_gradient_gemm_workspace_bytes': (c_size_t, []),
    'vipnerf_param_gradient_gemm': (c_int, [c_void_p, c_int32, c_int32, c_void_p, c_int32, c_int32, c_int64, c_void_p,
                                            c_int32, c_int32, c_void_p, c_int32, c_void_p, c_size_t, c_void_p]),
    'vipnerf_grad_scale': (c_float, [c_float]),
    'vipnerf_debug_set_profile_buffer': (c_int, [c_void_p]),
}

_lock = threading.Lock()
_lib = None


class VipNeRFLibraryError(RuntimeError):
    pass


def load() -> ctypes.CDLL:
    """Loads the in-tree shared library (once).  Raises if it has not been built."""
    global _lib
    with _lock:
        if _lib is not None:
            return _lib
        if not os.path.isfile(LIB_PATH):
            raise VipNeRFLibraryError(
                f'{LIB_PATH} is missing: the CUDA extension has not been built. '
                f'Run `python -m vipnerf_b200.build` (needs nvcc with sm_100a support). There is no CPU fallback.')
        lib = ctypes.CDLL(LIB_PATH)
        for name, (restype, argtypes) in EXPORTS.items():
            fn = getattr(lib, name)
            fn.restype = restype
            fn.argtypes = argtypes
        if lib.vipnerf_abi_version() != ABI_VERSION:
            raise VipNeRFLibraryError(f'ABI mismatch: library {lib.vipnerf_abi_version()}, binding {ABI_VERSION}')
        _lib = lib
        return lib


def check(status: int, what: str) -> None:
    """Maps the C status to the exception types the reference raises for the same situations
    (NotImplementedError for unsupported configurations, RuntimeError otherwise; VipNeRF01.py:292,311,321)."""
    if status == 0:
        return
    msg = load().vipnerf_last_error().decode(errors='replace')
    text = f'{what}: {STATUS_NAMES.get(status, status)}: {msg}'
    if status == -2:
        raise NotImplementedError(text)
    raise VipNeRFLibraryError(text)


def make_cfg(n_coarse=64, n_fine=128, n_sec_views=0, ndc=False, white_bkgd=False, lindisp=False, precision='bf16',
             l_pts=10, l_view=4, depth=8, width=256, skip=4, train_tf32=False, train_precision=None) -> Cfg:
    """train_precision: arithmetic of the training entry points - 'fp32' (CUDA cores), 'tf32' or 'fp16' (tensor cores);
    train_tf32=True is the older spelling of 'tf32'."""
    if train_precision is None:
        train_precision = 'tf32' if train_tf32 else 'fp32'
    if train_precision not in TRAIN_PRECISIONS:
        raise ValueError(f'train_precision = {train_precision!r}')
    flags = ((FLAG_NDC if ndc else 0) | (FLAG_WHITE_BKGD if white_bkgd else 0) | (FLAG_LINDISP if lindisp else 0)
             | (FLAG_TRAIN_TF32 if train_precision == 'tf32' else 0) | (FLAG_TRAIN_F16 if train_precision == 'fp16' else 0))
    return Cfg(ABI_VERSION, n_coarse, n_fine, l_pts, l_view, depth, width, skip, n_sec_views, flags,
               PRECISION[precision], 0)
