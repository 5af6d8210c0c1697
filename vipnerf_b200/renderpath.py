"""Torch-facing wrappers of the C ABI: tensors in, tensors out, everything on the caller's CUDA device and
current stream.  PyTorch is only plumbing here (device memory + streams); all arithmetic happens inside
libvipnerf_b200.so.  Function names follow the reference functions they stand in for
(src/models/VipNeRF01.py): render_rays :74, get_z_vals_coarse :173, volume_rendering :331,
get_z_vals_fine :205, MLP.forward :509.
"""
from __future__ import annotations

import ctypes
from typing import Dict, Iterable, Optional

import torch

from . import _lib

MLP_PARAM_ORDER = tuple(
    [f'pts_linears.{i}.{t}' for i in range(8) for t in ('weight', 'bias')]
    + [f'{name}.{t}' for name in ('views_linears.0', 'pts_output_linear', 'feature_linear', 'views_output_linear')
       for t in ('weight', 'bias')])

EVAL_KEYS = ('rgb', 'acc', 'alpha', 'depth', 'depth_var')
NDC_KEYS = ('depth_ndc', 'depth_var_ndc')
RAW_PER_SAMPLE_KEYS = ('z_vals', 'visibility', 'weights')


def _require_cuda(t: torch.Tensor, name: str) -> None:
    if not t.is_cuda:
        raise RuntimeError(f'{name} is on {t.device}: the ViP-NeRF B200 render path only runs on CUDA tensors '
                           f'(there is no CPU fallback)')


def _f32c(t: torch.Tensor, name: str) -> torch.Tensor:
    _require_cuda(t, name)
    if t.dtype != torch.float32:
        t = t.float()
    t = t.contiguous()
    if t.data_ptr() % 16 != 0:
        # a ray-dimension slice (batch[k][lo:hi] of sharding.shard_batch, Trainer01.py:83-87 sub-batches) is contiguous
        # but starts at 12*lo / 4*lo bytes; the library reads rays with 16-byte vector loads
        t = t.clone(memory_format=torch.contiguous_format)
    return t


def _ptr(t: Optional[torch.Tensor]) -> Optional[int]:
    return None if t is None else t.data_ptr()


def _stream(device) -> int:
    return torch.cuda.current_stream(device).cuda_stream


_linspace_cache: Dict[tuple, torch.Tensor] = {}


def _linspace(n: int, device) -> torch.Tensor:
    """torch.linspace(0, 1, n) evaluated on the CPU and copied, exactly as the reference does
    (VipNeRF01.py:186, :239): the table is bit-identical to the reference's."""
    key = (n, str(device))
    t = _linspace_cache.get(key)
    if t is None:
        t = torch.linspace(0., 1., steps=n).to(device)
        _linspace_cache[key] = t
    return t


def pack_mlp(params: Dict[str, torch.Tensor], precision: str, **cfg_kwargs) -> torch.Tensor:
    """Packs one MLP's 24 tensors (reference state_dict names, any prefix stripped) into the layout the
    kernels of `precision` read.  Returns a uint8 CUDA tensor."""
    lib = _lib.load()
    cfg = _lib.make_cfg(precision=precision, **cfg_kwargs)
    tensors = [_f32c(params[k], k) for k in MLP_PARAM_ORDER]
    device = tensors[0].device
    n_bytes = lib.vipnerf_packed_weight_bytes(ctypes.byref(cfg))
    if n_bytes == 0:
        _lib.check(lib.vipnerf_check_config(ctypes.byref(cfg)), 'vipnerf_check_config')
    packed = torch.zeros(n_bytes + 1024, dtype=torch.uint8, device=device)
    off = (-packed.data_ptr()) % 1024
    packed = packed[off:off + n_bytes]
    arr = (ctypes.c_void_p * 24)(*[t.data_ptr() for t in tensors])
    with torch.cuda.device(device):
        _lib.check(lib.vipnerf_pack_weights(ctypes.byref(cfg), arr, packed.data_ptr(), _stream(device)),
                   'vipnerf_pack_weights')
    packed._vipnerf_keepalive = tensors  # the pack kernel reads them asynchronously
    return packed


def _make_rays(batch: Dict[str, torch.Tensor], ndc: bool, n_coarse: int, n_fine: int, n_sec_views: int,
               keep: list) -> _lib.Rays:
    rays = _lib.Rays()
    device = batch['rays_o'].device
    names = ['rays_o', 'rays_d', 'view_dirs', 'near', 'far']
    if ndc:
        names += ['rays_o_ndc', 'rays_d_ndc', 'near_ndc', 'far_ndc']
    if n_sec_views > 0:
        names.append('rays_o2')
    for name in names + ['t_rand', 'u_rand']:
        if name in batch and batch[name] is not None:
            t = _f32c(batch[name], name)
            keep.append(t)
            setattr(rays, name, t.data_ptr())
    t_vals = _linspace(n_coarse, device)
    rays.t_vals = t_vals.data_ptr()
    if n_fine > 0:
        rays.u_vals = _linspace(n_fine, device).data_ptr()
    return rays


def _alloc_pass(out: _lib.PassOut, keys: Iterable[str], R: int, S: int, V: int, device,
                provided: Optional[Dict[str, torch.Tensor]] = None, tag: str = '') -> Dict[str, torch.Tensor]:
    """Output tensors of one pass.  `provided` maps reference-style names (`rgb_fine`, ...) to caller-owned fp32 CUDA
    tensors the kernel writes into instead of fresh allocations - e.g. views of another GPU's symmetric memory
    (sharding.PeerGather): the ray warps then store the finished maps directly over NVLink."""
    shapes = {'rgb': (R, 3), 'acc': (R,), 'depth': (R,), 'depth_var': (R,), 'depth_ndc': (R,), 'depth_var_ndc': (R,),
              'visibility2': (R, V), 'alpha': (R, S), 'z_vals': (R, S), 'visibility': (R, S), 'weights': (R, S),
              'raw_sigma': (R, S, 1), 'raw_rgb': (R, S, 3), 'raw_visibility': (R, S, 1),
              'raw_visibility2': (R, S, V, 1)}
    tensors, fresh, total = {}, [], 0
    for k in keys:
        t = provided.get(f'{k}_{tag}') if provided else None
        if t is None:
            n = 1
            for d in shapes[k]:
                n *= int(d)
            fresh.append((k, total, n))
            total += (n + 3) // 4 * 4          # every array starts on a 16-byte boundary
            continue
        if (tuple(t.shape) != tuple(shapes[k]) or t.dtype != torch.float32 or not t.is_cuda or not t.is_contiguous()
                or t.data_ptr() % 4 != 0):
            raise ValueError(f'out[{k}_{tag}]: expected a contiguous fp32 CUDA tensor of shape {tuple(shapes[k])}, got '
                             f'{tuple(t.shape)} {t.dtype} on {t.device}')
        tensors[k] = t
    if fresh:
        # ONE allocation per pass (the arrays are views of it): 14 allocator calls per eval render were ~40 us of host time
        flat = torch.empty(max(total, 4), dtype=torch.float32, device=device)
        for k, off, n in fresh:
            tensors[k] = flat[off:off + n].view(shapes[k])
    for k in keys:
        setattr(out, k, tensors[k].data_ptr())
    return {k: tensors[k] for k in keys}


_workspace_cache: Dict[tuple, torch.Tensor] = {}


def _workspace(n_bytes: int, device) -> torch.Tensor:
    key = (str(device), torch.cuda.current_stream(device).cuda_stream)
    ws = _workspace_cache.get(key)
    if ws is None or ws.numel() < n_bytes + 256:
        ws = torch.empty(n_bytes + 256, dtype=torch.uint8, device=device)
        _workspace_cache[key] = ws
    off = (-ws.data_ptr()) % 256
    return ws[off:off + n_bytes]


def pass_keys(ndc: bool, retraw: bool, n_sec_views: int) -> list:
    keys = list(EVAL_KEYS) + (list(NDC_KEYS) if ndc else [])
    if n_sec_views > 0:
        keys.append('visibility2')
    if retraw:
        keys += list(RAW_PER_SAMPLE_KEYS) + ['raw_sigma', 'raw_rgb', 'raw_visibility']
        if n_sec_views > 0:
            keys.append('raw_visibility2')
    return keys


def render_rays(batch: Dict[str, torch.Tensor], packed_coarse: torch.Tensor, packed_fine: Optional[torch.Tensor], *,
                ndc: bool, precision: str, n_coarse: int = 64, n_fine: int = 128, retraw: bool = False,
                n_sec_views: int = 0, white_bkgd: bool = False, lindisp: bool = False,
                keys: Optional[Iterable[str]] = None,
                out_tensors: Optional[Dict[str, torch.Tensor]] = None) -> Dict[str, torch.Tensor]:
    """VipNeRF.render_rays (VipNeRF01.py:74-171) for every ray of `batch` in one library call.
    Returns the reference's output dict (`<key>_coarse` / `<key>_fine`); `out_tensors` (optional) supplies the storage
    of some of them."""
    lib = _lib.load()
    rays_o = batch['rays_o']
    _require_cuda(rays_o, 'rays_o')
    device = rays_o.device
    R = rays_o.shape[0]
    has_fine = packed_fine is not None and n_fine > 0
    cfg = _lib.make_cfg(n_coarse=n_coarse, n_fine=n_fine if has_fine else 0, n_sec_views=n_sec_views, ndc=ndc,
                        white_bkgd=white_bkgd, lindisp=lindisp, precision=precision)
    keep = []
    rays = _make_rays(batch, ndc, n_coarse, cfg.n_fine, n_sec_views, keep)
    out = _lib.Out()
    wanted = list(keys) if keys is not None else pass_keys(ndc, retraw, n_sec_views)
    result = {}
    tensors_c = _alloc_pass(out.coarse, wanted, R, n_coarse, n_sec_views, device, out_tensors, 'coarse')
    result.update({f'{k}_coarse': v for k, v in tensors_c.items()})
    if has_fine:
        tensors_f = _alloc_pass(out.fine, wanted, R, n_coarse + n_fine, n_sec_views, device, out_tensors, 'fine')
        result.update({f'{k}_fine': v for k, v in tensors_f.items()})
    with torch.cuda.device(device):
        ws_bytes = lib.vipnerf_workspace_bytes(ctypes.byref(cfg), R)
        if ws_bytes == 0:
            _lib.check(lib.vipnerf_check_config(ctypes.byref(cfg)), 'vipnerf_check_config')
        ws = _workspace(ws_bytes, device)
        _lib.check(lib.vipnerf_render_forward(ctypes.byref(cfg), ctypes.byref(rays), R, packed_coarse.data_ptr(),
                                              packed_fine.data_ptr() if has_fine else None, ctypes.byref(out),
                                              ws.data_ptr(), ws_bytes, _stream(device)), 'vipnerf_render_forward')
    for tag in ('coarse', 'fine'):
        if f'raw_rgb_{tag}' in result:  # the reference returns the same tensor under both names (:531)
            result[f'raw_rgb_view_dependent_{tag}'] = result[f'raw_rgb_{tag}']
    return result


def coarse_z_vals(batch: Dict[str, torch.Tensor], *, ndc: bool, n_coarse: int = 64, lindisp: bool = False,
                  precision: str = 'fp32') -> torch.Tensor:
    """get_z_vals_coarse (VipNeRF01.py:173-203)."""
    lib = _lib.load()
    device = batch['rays_o'].device
    R = batch['rays_o'].shape[0]
    cfg = _lib.make_cfg(n_coarse=n_coarse, n_fine=0, ndc=ndc, lindisp=lindisp, precision=precision)
    keep = []
    rays = _make_rays(batch, ndc, n_coarse, 0, 0, keep)
    z = torch.empty((R, n_coarse), dtype=torch.float32, device=device)
    with torch.cuda.device(device):
        _lib.check(lib.vipnerf_coarse_z(ctypes.byref(cfg), ctypes.byref(rays), R, z.data_ptr(), _stream(device)),
                   'vipnerf_coarse_z')
    return z


def mlp_forward(batch: Dict[str, torch.Tensor], z_vals: torch.Tensor, packed: torch.Tensor, *, ndc: bool,
                precision: str, n_sec_views: int = 0) -> Dict[str, torch.Tensor]:
    """MLP.forward on the sample points pts_o + pts_d * z (VipNeRF01.py:105-107, :509-535).
    Returns sigma [R,S,1], rgb [R,S,3], visibility [R,S,1] (+ visibility2 [R,S,V,1])."""
    lib = _lib.load()
    z_vals = _f32c(z_vals, 'z_vals')
    device = z_vals.device
    R, S = z_vals.shape
    cfg = _lib.make_cfg(n_coarse=64, n_fine=0, n_sec_views=n_sec_views, ndc=ndc, precision=precision)
    keep = []
    rays = _make_rays(batch, ndc, 64, 0, n_sec_views, keep)
    out = _lib.PassOut()
    keys = ['raw_sigma', 'raw_rgb', 'raw_visibility'] + (['raw_visibility2'] if n_sec_views else [])
    t = _alloc_pass(out, keys, R, S, n_sec_views, device)
    with torch.cuda.device(device):
        _lib.check(lib.vipnerf_mlp_forward(ctypes.byref(cfg), ctypes.byref(rays), R, S, z_vals.data_ptr(),
                                           packed.data_ptr(), ctypes.byref(out), None, 0, _stream(device)),
                   'vipnerf_mlp_forward')
    res = {'sigma': t['raw_sigma'], 'rgb': t['raw_rgb'], 'visibility': t['raw_visibility']}
    if n_sec_views:
        res['visibility2'] = t['raw_visibility2']
    return res


def volume_rendering(batch: Dict[str, torch.Tensor], z_vals: torch.Tensor, sigma: torch.Tensor, rgb: torch.Tensor,
                     visibility2: Optional[torch.Tensor] = None, *, ndc: bool, white_bkgd: bool = False,
                     n_fine: int = 0, u_rand: Optional[torch.Tensor] = None) -> Dict[str, torch.Tensor]:
    """volume_rendering (VipNeRF01.py:331-384) on given network outputs; with n_fine > 0 also returns
    'z_vals_fine' = get_z_vals_fine (:205-216) computed from the composited weights."""
    lib = _lib.load()
    z_vals = _f32c(z_vals, 'z_vals')
    device = z_vals.device
    R, S = z_vals.shape
    sigma = _f32c(sigma, 'sigma').reshape(R, S)
    rgb = _f32c(rgb, 'rgb').reshape(R, S, 3)
    V = 0
    if visibility2 is not None:
        visibility2 = _f32c(visibility2, 'visibility2').reshape(R, S, -1)
        V = visibility2.shape[-1]
    cfg = _lib.make_cfg(n_coarse=S, n_fine=n_fine, n_sec_views=V, ndc=ndc, white_bkgd=white_bkgd, precision='fp32')
    keep = []
    b = dict(batch)
    if u_rand is not None:
        b['u_rand'] = u_rand
    rays = _make_rays(b, ndc, S, n_fine, 0, keep)
    out = _lib.PassOut()
    keys = pass_keys(ndc, True, V)
    keys = [k for k in keys if not k.startswith('raw_') and k != 'z_vals']
    t = _alloc_pass(out, keys, R, S, V, device)
    z_fine = torch.empty((R, S + n_fine), dtype=torch.float32, device=device) if n_fine > 0 else None
    with torch.cuda.device(device):
        _lib.check(lib.vipnerf_composite(ctypes.byref(cfg), ctypes.byref(rays), R, S, z_vals.data_ptr(),
                                         sigma.data_ptr(), rgb.data_ptr(), _ptr(visibility2), ctypes.byref(out),
                                         _ptr(z_fine), _stream(device)), 'vipnerf_composite')
    if z_fine is not None:
        t['z_vals_fine'] = z_fine
    return t
