"""CPU oracle for the ViP-NeRF volumetric render path.  TEST INFRASTRUCTURE ONLY.

This file is a from-scratch CPU (torch, fp32) restatement of the algorithm in the
reference's ``src/models/VipNeRF01.py``.  It exists so that the CUDA path in
``vipnerf_b200/`` can be checked on machines where ``/root/reference`` is absent
(the GPU box).  Only ``tests/``, ``__graft_entry__.smoke()`` and the
``cpu_baseline`` / ``--impl reference`` legs of ``bench.py`` may import it; the
product package never does.

Parity status: PINNED.  The reference ships no tests or golden vectors for this
path, so the oracle is pinned the other way the brief
allows: ``oracle/make_golden.py`` imports the unmodified reference in the build
container, runs it on seeded inputs and commits the outputs under
``tests/golden/``; ``tests/test_oracle_golden.py`` checks this restatement
against those vectors (and against the live reference when it is mounted).

Every function cites the reference lines it follows (paths are relative to
``/root/reference/src/models/VipNeRF01.py`` unless stated).

Data conventions: ``R`` rays, ``S`` samples per ray, all tensors fp32 and
row-major; an "MLP parameter dict" has the reference's ``state_dict`` keys of one
``MLP`` (``pts_linears.{0..7}.{weight,bias}``, ``views_linears.0.*``,
``pts_output_linear.*``, ``feature_linear.*``, ``views_output_linear.*``).
"""
from __future__ import annotations

import math
from typing import Dict, Optional

import numpy
import torch
import torch.nn.functional as F

# --------------------------------------------------------------------------------------
# Deterministic synthetic weights / rays (shared by golden generation, tests, bench, smoke)
# --------------------------------------------------------------------------------------

MLP_SHAPES = (
    # key, (out, in)  -- reference MLP.__init__ :472-491 for D=8, W=256, L_pts=10, L_view=4
    ('pts_linears.0', (256, 63)),
    ('pts_linears.1', (256, 256)),
    ('pts_linears.2', (256, 256)),
    ('pts_linears.3', (256, 256)),
    ('pts_linears.4', (256, 256)),
    ('pts_linears.5', (256, 319)),
    ('pts_linears.6', (256, 256)),
    ('pts_linears.7', (256, 256)),
    ('views_linears.0', (128, 283)),
    ('pts_output_linear', (1, 256)),
    ('feature_linear', (256, 256)),
    ('views_output_linear', (4, 128)),
)


def _splitmix_uniform(n: int, seed: int) -> numpy.ndarray:
    """n floats in [0,1) from a splitmix64 counter hash: pure integer numpy arithmetic, so the
    stream is identical on every machine / numpy / torch version (unlike torch.manual_seed)."""
    with numpy.errstate(over='ignore'):
        x = numpy.arange(n, dtype=numpy.uint64) + numpy.uint64((seed * 0x9E3779B97F4A7C15) & 0xFFFFFFFFFFFFFFFF)
        x = x + numpy.uint64(0x9E3779B97F4A7C15)
        x = (x ^ (x >> numpy.uint64(30))) * numpy.uint64(0xBF58476D1CE4E5B9)
        x = (x ^ (x >> numpy.uint64(27))) * numpy.uint64(0x94D049BB133111EB)
        x = x ^ (x >> numpy.uint64(31))
    return ((x >> numpy.uint64(40)).astype(numpy.float64) * (1.0 / (1 << 24))).astype(numpy.float32)


def synth_mlp_params(seed: int, sigma_gain: float = 300.0, sigma_bias: float = 1.0) -> Dict[str, torch.Tensor]:
    """Well-conditioned synthetic weights of one MLP: uniform(+-1/sqrt(fan_in)) like torch.nn.Linear's
    default init, with the density head rescaled so that rays saturate (acc ~ 1).  Default-initialised
    weights give sigma ~ 0 and make depth / visibility2 ill-conditioned."""
    params = {}
    for i, (key, (n_out, n_in)) in enumerate(MLP_SHAPES):
        bound = 1.0 / math.sqrt(n_in)
        w = (_splitmix_uniform(n_out * n_in, seed * 1000 + 2 * i) * 2 - 1) * bound
        b = (_splitmix_uniform(n_out, seed * 1000 + 2 * i + 1) * 2 - 1) * bound
        params[f'{key}.weight'] = torch.from_numpy(w.reshape(n_out, n_in).copy())
        params[f'{key}.bias'] = torch.from_numpy(b.copy())
    params['pts_output_linear.weight'] = params['pts_output_linear.weight'] * sigma_gain
    params['pts_output_linear.bias'] = torch.full((1,), float(sigma_bias))
    return params


def synth_state_dict(seed: int = 0) -> Dict[str, torch.Tensor]:
    """Full-model state dict with the reference's key names (`coarse_model.*`, `fine_model.*`)."""
    sd = {}
    # per-MLP seeds screened so that both networks give acc ~ 1 on the LLFF (NDC) and DTU ray fixtures
    good = (1, 3, 7, 8, 9, 4)
    for prefix, s in (('coarse_model', good[(2 * seed) % 6]), ('fine_model', good[(2 * seed + 1) % 6])):
        for k, v in synth_mlp_params(s).items():
            sd[f'{prefix}.{k}'] = v
    return sd


def split_state_dict(sd: Dict[str, torch.Tensor], prefix: str) -> Dict[str, torch.Tensor]:
    plen = len(prefix) + 1
    return {k[plen:]: v for k, v in sd.items() if k.startswith(prefix + '.')}


# Camera models of the BASELINE configs (values from the reference's committed
# runs/training/train00{12,01,42}/<scene>/ModelConfigs.json).
SCENES = {
    'fern': dict(ndc=True, h=756, w=1008, f=815.1315832201474, near=1.0, far=6.179403816938637),
    'fern_half': dict(ndc=True, h=378, w=504, f=407.5657916100737, near=1.0, far=6.179403816938637),
    're10k': dict(ndc=True, h=576, w=1024, f=493.91, near=1.0, far=133.3),
    'dtu': dict(ndc=False, h=300, w=400, f=721.94, near=0.09, far=5.0),
    'synthetic32': dict(ndc=False, h=32, w=32, f=40.0, near=0.09, far=5.0),
}


def _pose_from_seed(seed: int) -> numpy.ndarray:
    """A small camera motion around the identity (camera-to-world 3x4, NeRF/LLFF axes: x right, y up,
    looking down -z).  Rotations of a few degrees and a translation of ~0.1 scene units - the scale of the
    reference's spiral video poses after `translation_scale`."""
    u = _splitmix_uniform(6, 7000 + seed).astype(numpy.float64) * 2 - 1
    ax, ay, az = u[:3] * 0.06
    rx = numpy.array([[1, 0, 0], [0, math.cos(ax), -math.sin(ax)], [0, math.sin(ax), math.cos(ax)]])
    ry = numpy.array([[math.cos(ay), 0, math.sin(ay)], [0, 1, 0], [-math.sin(ay), 0, math.cos(ay)]])
    rz = numpy.array([[math.cos(az), -math.sin(az), 0], [math.sin(az), math.cos(az), 0], [0, 0, 1]])
    pose = numpy.zeros((3, 4))
    pose[:, :3] = rz @ ry @ rx
    pose[:, 3] = u[3:] * 0.12
    return pose.astype(numpy.float32)


def make_rays(scene: str, n_rays: int, seed: int = 1, n_sec_views: int = 0, first_pixel: Optional[int] = None
              ) -> Dict[str, torch.Tensor]:
    """Ray batch dict with the keys `VipNeRF.forward` consumes, for `n_rays`
    consecutive pixels of a synthetic camera with the scene's real intrinsics.

    Follows DataPreprocessor01.get_rays :335-352 (pinhole, y/z flipped, rotate by c2w),
    get_ndc_rays :355-373, get_view_dirs :376-378 and create_test_data :821-854 for the dict layout.
    The arithmetic is restated in numpy float32 - it only has to produce *realistic* inputs; both the
    oracle and the CUDA path consume the same tensors."""
    sc = SCENES[scene]
    h, w, f = sc['h'], sc['w'], numpy.float32(sc['f'])
    pose = _pose_from_seed(seed)
    if first_pixel is None:
        first_pixel = int(_splitmix_uniform(1, 9000 + seed)[0] * max(1, h * w - n_rays))
    pix = (numpy.arange(n_rays, dtype=numpy.int64) + first_pixel) % (h * w)
    px = (pix % w).astype(numpy.float32)
    py = (pix // w).astype(numpy.float32)
    cx, cy = numpy.float32(w / 2), numpy.float32(h / 2)
    dirs = numpy.stack([(px - cx) / f, -(py - cy) / f, -numpy.ones_like(px)], axis=-1)  # camera frame
    rays_d = (dirs[:, None, :] * pose[None, :3, :3]).sum(-1).astype(numpy.float32)
    rays_o = numpy.broadcast_to(pose[:3, 3], rays_d.shape).astype(numpy.float32).copy()
    view_dirs = (rays_d / numpy.linalg.norm(rays_d, axis=-1, keepdims=True)).astype(numpy.float32)
    batch = {
        'rays_o': torch.from_numpy(rays_o),
        'rays_d': torch.from_numpy(rays_d),
        'view_dirs': torch.from_numpy(view_dirs),
        'near': torch.full((n_rays, 1), float(sc['near'])),
        'far': torch.full((n_rays, 1), float(sc['far'])),
    }
    if sc['ndc']:
        near = numpy.float32(sc['near'])
        t = -(near + rays_o[:, 2]) / rays_d[:, 2]
        o = rays_o + t[:, None] * rays_d
        sx, sy = numpy.float32(-1. / (w / (2. * f))), numpy.float32(-1. / (h / (2. * f)))
        o_ndc = numpy.stack([sx * o[:, 0] / o[:, 2], sy * o[:, 1] / o[:, 2], 1. + 2. * near / o[:, 2]], -1)
        d_ndc = numpy.stack([sx * (rays_d[:, 0] / rays_d[:, 2] - o[:, 0] / o[:, 2]),
                             sy * (rays_d[:, 1] / rays_d[:, 2] - o[:, 1] / o[:, 2]),
                             -2. * near / o[:, 2]], -1)
        batch['rays_o_ndc'] = torch.from_numpy(o_ndc.astype(numpy.float32))
        batch['rays_d_ndc'] = torch.from_numpy(d_ndc.astype(numpy.float32))
        batch['near_ndc'] = torch.zeros((n_rays, 1))
        batch['far_ndc'] = torch.ones((n_rays, 1))
    if n_sec_views > 0:
        centres = [_pose_from_seed(seed * 31 + 17 * (v + 1))[:3, 3] * 2.0 for v in range(n_sec_views)]
        o2 = numpy.broadcast_to(numpy.stack(centres, 0)[None], (n_rays, n_sec_views, 3))
        batch['rays_o2'] = torch.from_numpy(o2.astype(numpy.float32).copy())
    return batch


def make_supervision(scene: str, n_rays: int, n_sec: int, iter_num: int = 40000) -> dict:
    """Synthetic supervision with the keys the reference's losses read (MSE01.py:30-31, SparseDepthMSE01.py:33-37,
    VisibilityPriorLoss01.py:33-36): two thirds of the rays carry colour targets, the rest sparse depths."""
    sc = SCENES[scene]
    u = _splitmix_uniform(n_rays * 8, 4242).reshape(n_rays, 8)
    mask_nerf = numpy.arange(n_rays) < (2 * n_rays) // 3
    return {
        'target_rgb': torch.from_numpy(u[:, :3].copy()),
        'indices_mask_nerf': torch.from_numpy(mask_nerf),
        'indices_mask_sparse_depth': torch.from_numpy(~mask_nerf),
        'sparse_depth_values': torch.from_numpy((sc['near'] + u[:, 3:4] * (min(sc['far'], 8.0) - sc['near'])).astype('float32')),
        'visibility_prior_masks': torch.from_numpy((u[:, 4:4 + n_sec] > 0.3).astype('float32')),
        'iter_num': iter_num,
        'num_frames': n_sec + 1,
    }


# --------------------------------------------------------------------------------------
# Stage functions
# --------------------------------------------------------------------------------------

def positional_encoding(x: torch.Tensor, degree: int) -> torch.Tensor:
    """[..., 3] -> [..., 3 + 6*degree]: (x, sin(2^0 x), cos(2^0 x), ..., sin(2^(L-1) x), cos(2^(L-1) x)).
    Reference: PositionalEncoder.create_pos_enc_fns/encode :424-448 with the kwargs of
    MLP.get_positional_encoder :494-507 (include_input, log-sampled powers of two)."""
    parts = [x]
    for k in range(degree):
        xf = x * float(2 ** k)
        parts.append(torch.sin(xf))
        parts.append(torch.cos(xf))
    return torch.cat(parts, dim=-1)


def _round_bf16(t: torch.Tensor) -> torch.Tensor:
    return t.to(torch.bfloat16).to(torch.float32)


def _round_fp16(t: torch.Tensor) -> torch.Tensor:
    return t.clamp(-65504.0, 65504.0).to(torch.float16).to(torch.float32)


def _round_tf32(t: torch.Tensor) -> torch.Tensor:
    bits = t.contiguous().view(torch.int32)
    return ((bits + 0x1000) & ~0x1FFF).view(torch.float32)


def linear(x: torch.Tensor, w: torch.Tensor, b: Optional[torch.Tensor], mode: str = 'fp32') -> torch.Tensor:
    """y = x W^T + b.  `mode` emulates the arithmetic of the CUDA kernels' tensor-core layers:
    'fp32' (the reference), 'bf16' (both operands rounded to bf16, fp32 accumulate) and
    'bf16x3' (hi/lo split of both operands, the lo*lo term dropped)."""
    if mode == 'fp32':
        return F.linear(x, w, b)
    if mode == 'bf16':
        # the kernel adds the bias on the tensor core too (as a bf16 weight column against a constant-one input)
        return F.linear(_round_bf16(x), _round_bf16(w), _round_bf16(b) if b is not None else None)
    if mode == 'bf16x3':
        xh, wh = _round_bf16(x), _round_bf16(w)
        xl, wl = _round_bf16(x - xh), _round_bf16(w - wh)
        return F.linear(xh, wh, b) + (F.linear(xl, wh) + F.linear(xh, wl))
    # modes below exist for tools/precision_study.py (error of candidate tensor-core arithmetics, measured on the CPU)
    if mode == 'fp16':      # kind::f16 with fp16 operands: 11-bit significands at the bf16 MMA rate, fp32 accumulate
        return F.linear(_round_fp16(x), _round_fp16(w), _round_fp16(b) if b is not None else None)
    if mode == 'fp16_a2':   # 2 MMAs: activations split hi + lo (fp16), weights rounded once
        xh, wh = _round_fp16(x), _round_fp16(w)
        return F.linear(xh, wh, _round_fp16(b) if b is not None else None) + F.linear(_round_fp16(x - xh), wh)
    if mode == 'bf16_a2':   # 2 MMAs: activations split hi + lo (bf16), weights rounded once
        xh, wh = _round_bf16(x), _round_bf16(w)
        return F.linear(xh, wh, _round_bf16(b) if b is not None else None) + F.linear(_round_bf16(x - xh), wh)
    if mode == 'tf32':      # kind::tf32: 11-bit significands at half the bf16 rate
        return F.linear(_round_tf32(x), _round_tf32(w), b)
    raise ValueError(mode)


def mlp_forward(params: Dict[str, torch.Tensor], pts: torch.Tensor, view_dirs: torch.Tensor,
                view_dirs2: Optional[torch.Tensor] = None, l_pts: int = 10, l_view: int = 4,
                mode: str = 'fp32', sigma_noise: Optional[torch.Tensor] = None) -> Dict[str, torch.Tensor]:
    """One radiance/visibility MLP on P points.  pts [P,3], view_dirs [P,3], view_dirs2 [P,V,3].

    Reference: MLP.forward :509-535, get_view_independent_outputs :537-566 (8 ReLU layers, the skip
    concatenates *encoded input first* after layer 4 :543-544, sigma = relu(head (+noise)) :549-554,
    feature has no activation :564), get_view_dependent_outputs :568-596 (cat([feature, enc(view)]) :576,
    283->128 ReLU, 128->4, sigmoid on rgb[0:3] and visibility[3]).

    In the tensor-core modes only the layers the CUDA kernel runs on tensor cores are emulated at reduced
    precision (the eight trunk layers, feature_linear and the feature columns of views_linears.0); the
    density head, the PRIMARY view-direction columns of views_linears.0 and views_output_linear stay fp32, as in
    the kernel (the secondary views' direction columns are a tensor-core step there)."""
    enc = positional_encoding(pts, l_pts)
    h = enc
    for i in range(8):
        h = F.relu(linear(h, params[f'pts_linears.{i}.weight'], params[f'pts_linears.{i}.bias'], mode))
        if i == 4:
            h = torch.cat([enc, h], dim=-1)
    sigma_raw = F.linear(h, params['pts_output_linear.weight'], params['pts_output_linear.bias'])
    if sigma_noise is not None:
        sigma_raw = sigma_raw + sigma_noise
    sigma = F.relu(sigma_raw)
    wv, bv = params['views_linears.0.weight'], params['views_linears.0.bias']
    wo, bo = params['views_output_linear.weight'], params['views_output_linear.bias']
    n_feat = params['feature_linear.weight'].shape[0]
    if mode == 'fp32':
        feature = F.linear(h, params['feature_linear.weight'], params['feature_linear.bias'])
        feat_part = None
    else:
        # the tensor-core kernels fold feature_linear (no activation, :564) into the feature columns of views_linears.0:
        # Wv_f (W8 h + b8) = (Wv_f W8) h + Wv_f b8, the product matrix formed in double precision at pack time
        # (vipnerf_b200/csrc/layout.cuh); `feature` itself is not an output (:533-534)
        wf = (wv[:, :n_feat].double() @ params['feature_linear.weight'].double()).float()
        bv = (bv.double() + wv[:, :n_feat].double() @ params['feature_linear.bias'].double()).float()
        feature = None
        feat_part = linear(h, wf, None, mode)

    def view_head(enc_view, feat, feat_pre, secondary=False):
        if mode == 'fp32':
            hv = F.relu(F.linear(torch.cat([feat, enc_view], dim=-1), wv, bv))
        elif secondary:
            # secondary views: the direction differs per sample, so the kernel runs these 27 columns (and the bias)
            # on the tensor core as well
            hv = F.relu(feat_pre + linear(enc_view, wv[:, n_feat:], bv, mode))
        else:
            hv = F.relu(feat_pre + F.linear(enc_view, wv[:, n_feat:], bv))
        return torch.sigmoid(F.linear(hv, wo, bo))

    out = view_head(positional_encoding(view_dirs, l_view), feature, feat_part)
    result = {'sigma': sigma, 'rgb': out[..., :3], 'visibility': out[..., 3:4]}
    if view_dirs2 is not None:
        v = view_dirs2.shape[1]
        feat_v = feature[:, None, :].expand(-1, v, -1) if feature is not None else None
        feat_pre_v = feat_part[:, None, :].expand(-1, v, -1) if feat_part is not None else None
        out2 = view_head(positional_encoding(view_dirs2, l_view), feat_v, feat_pre_v, secondary=True)
        result['visibility2'] = out2[..., 3:4]
    return result


def coarse_z_vals(near: torch.Tensor, far: torch.Tensor, n_samples: int, lindisp: bool = False,
                  t_rand: Optional[torch.Tensor] = None) -> torch.Tensor:
    """[R,1] near/far -> [R,n_samples].  Reference: get_z_vals_coarse :173-203 (linspace lerp, optional
    lindisp, stratified jitter between mid-points when `t_rand` [R,n] is given, i.e. train+perturb)."""
    t = torch.linspace(0., 1., steps=n_samples)
    if not lindisp:
        z = near * (1. - t) + far * t
    else:
        z = 1. / (1. / near * (1. - t) + 1. / far * t)
    z = z.expand(near.shape[0], n_samples)
    if t_rand is not None:
        mids = .5 * (z[..., 1:] + z[..., :-1])
        upper = torch.cat([mids, z[..., -1:]], -1)
        lower = torch.cat([z[..., :1], mids], -1)
        z = lower + (upper - lower) * t_rand
    return z


def sample_pdf(bins: torch.Tensor, weights: torch.Tensor, n_samples: int,
               u: Optional[torch.Tensor] = None) -> torch.Tensor:
    """Inverse-CDF sampling.  bins [R,B], weights [R,B-1] -> [R,n_samples].
    Reference: VipNeRF.sample_pdf :229-262 (+1e-5, cdf with leading 0, searchsorted(right=True), clamp of
    the bracketing indices, denom<1e-5 -> 1).  `u` None = deterministic linspace (eval)."""
    weights = weights + 1e-5
    pdf = weights / torch.sum(weights, -1, keepdim=True)
    cdf = torch.cumsum(pdf, -1)
    cdf = torch.cat([torch.zeros_like(cdf[..., :1]), cdf], -1)
    if u is None:
        u = torch.linspace(0., 1., steps=n_samples).expand(cdf.shape[0], n_samples)
    u = u.contiguous()
    idx = torch.searchsorted(cdf, u, right=True)
    below = torch.clamp(idx - 1, min=0)
    above = torch.clamp(idx, max=cdf.shape[-1] - 1)
    cdf_b, cdf_a = torch.gather(cdf, 1, below), torch.gather(cdf, 1, above)
    bin_b, bin_a = torch.gather(bins, 1, below), torch.gather(bins, 1, above)
    denom = cdf_a - cdf_b
    denom = torch.where(denom < 1e-5, torch.ones_like(denom), denom)
    t = (u - cdf_b) / denom
    return bin_b + t * (bin_a - bin_b)


def fine_z_vals(z_coarse: torch.Tensor, weights_coarse: torch.Tensor, n_fine: int,
                u: Optional[torch.Tensor] = None) -> torch.Tensor:
    """[R,Nc] -> [R,Nc+n_fine] sorted.  Reference: get_z_vals_fine :205-216 (mid-points as bins, the first
    and last coarse weight dropped, sort of the concatenation)."""
    mids = .5 * (z_coarse[..., 1:] + z_coarse[..., :-1])
    samples = sample_pdf(mids, weights_coarse[..., 1:-1], n_fine, u).detach()   # :213
    z, _ = torch.sort(torch.cat([z_coarse, samples], -1), -1)
    return z


def depth_from_ndc(z_ndc: torch.Tensor, rays_o: torch.Tensor, rays_d: torch.Tensor) -> torch.Tensor:
    """Reference: convert_depth_from_ndc :386-403 (near hard-coded to 1; +1e-3 only where z_ndc == 1)."""
    oz, dz = rays_o[..., 2:3], rays_d[..., 2:3]
    tn = -(1 + oz) / dz
    c = torch.where(z_ndc == 1., 1e-3, 0.)
    return (oz + tn * dz) / dz * (1 / (1 - z_ndc + c) - 1) + tn


def other_view_dirs(z_vals: torch.Tensor, rays_o: torch.Tensor, rays_d: torch.Tensor, rays_o2: torch.Tensor,
                    ndc: bool) -> torch.Tensor:
    """[R,S] -> [R,S,V,3] unit vectors from each secondary camera centre to each sample.
    Reference: compute_other_view_dirs :218-226 (note the NDC un-projection uses +1e-6, not the +1e-3
    rule of convert_depth_from_ndc)."""
    if ndc:
        tn = -(1 + rays_o[..., 2]) / rays_d[..., 2]
        z_vals = (((rays_o[..., None, 2] + tn[..., None] * rays_d[..., None, 2]) / (1 - z_vals + 1e-6))
                  - rays_o[..., None, 2]) / rays_d[..., None, 2]
    pts = rays_o[..., None, :] + z_vals[..., None] * rays_d[..., None, :]
    d = pts[:, :, None] - rays_o2[..., None, :, :]
    return d / torch.norm(d, dim=-1, keepdim=True)


def composite(sigma: torch.Tensor, rgb: torch.Tensor, z_vals: torch.Tensor, ray_dir_for_delta: torch.Tensor,
              ndc: bool, rays_o: Optional[torch.Tensor] = None, rays_d: Optional[torch.Tensor] = None,
              white_bkgd: bool = False, visibility2: Optional[torch.Tensor] = None) -> Dict[str, torch.Tensor]:
    """Alpha compositing of one sample set.  sigma [R,S], rgb [R,S,3], z_vals [R,S] (NDC z when `ndc`),
    `ray_dir_for_delta` = rays_d (world) or rays_d_ndc; visibility2 [R,S,V].
    Reference: volume_rendering :331-384 (last interval 1e10 world / 1-z NDC, alpha = 1-exp(-sigma*delta),
    transmittance = exclusive cumprod of (1-alpha+1e-10), depth = sum(w z)/(acc+1e-6), un-normalised
    depth_var, NDC depth conversion, optional white background, visibility2 = sum(w v2)/(acc+1e-6))."""
    last = torch.full_like(z_vals[..., :1], 1.0 if ndc else 1e10)
    dists = torch.cat([z_vals, last], -1)
    delta = (dists[..., 1:] - dists[..., :-1]) * torch.norm(ray_dir_for_delta[..., None, :], dim=-1)
    alpha = 1. - torch.exp(-sigma * delta)
    trans = torch.cumprod(torch.cat([torch.ones_like(alpha[:, :1]), 1. - alpha + 1e-10], -1), -1)[:, :-1]
    weights = alpha * trans
    rgb_map = torch.sum(weights[..., None] * rgb, dim=-2)
    acc = torch.sum(weights, dim=-1)
    out = {}

    def depth_stats(z):
        d = torch.sum(weights * z, dim=-1) / (acc + 1e-6)
        return d, torch.sum(weights * torch.square(z - d[..., None]), dim=-1)

    if ndc:
        out['depth_ndc'], out['depth_var_ndc'] = depth_stats(z_vals)
        out['depth'], out['depth_var'] = depth_stats(depth_from_ndc(z_vals, rays_o, rays_d))
    else:
        out['depth'], out['depth_var'] = depth_stats(z_vals)
    if white_bkgd:
        rgb_map = rgb_map + (1. - acc[..., None])
    out.update({'rgb': rgb_map, 'acc': acc, 'alpha': alpha, 'visibility': trans, 'weights': weights})
    if visibility2 is not None:
        out['visibility2'] = torch.sum(weights[..., None] * visibility2, dim=-2) / (acc[..., None] + 1e-6)
    return out


# --------------------------------------------------------------------------------------
# Whole path
# --------------------------------------------------------------------------------------

def draw_training_randoms(n_rays: int, n_coarse: int = 64, n_fine: int = 128, chunk: int = 4096,
                          netchunk: int = 16384, perturb: bool = True, raw_noise_std: float = 1.0,
                          has_fine: bool = True) -> Dict[str, torch.Tensor]:
    """Consumes torch's global CPU generator in exactly the order the reference's train-mode forward does and
    returns the draws as whole-batch tensors: per `chunk` of rays (batchify_rays :54) first torch.rand [r,Nc]
    (get_z_vals_coarse :200), then one torch.randn [n,1] per `netchunk` slice of the flattened coarse points
    (batchify :305 -> get_view_independent_outputs :551), then torch.rand [r,Nf] (sample_pdf :242) and the
    fine network's randn slices.  Keys: t_rand [R,Nc], u_rand [R,Nf], sigma_noise_coarse [R,Nc],
    sigma_noise_fine [R,Nc+Nf] (already multiplied by raw_noise_std); a key is absent when its source is off."""
    parts = {'t_rand': [], 'u_rand': [], 'sigma_noise_coarse': [], 'sigma_noise_fine': []}

    def noise(n_points, r, s):
        pieces = [torch.randn(min(netchunk, n_points - i), 1) * raw_noise_std for i in range(0, n_points, netchunk)]
        return torch.cat(pieces, 0).reshape(r, s)

    for i in range(0, n_rays, chunk):
        r = min(chunk, n_rays - i)
        if perturb:
            parts['t_rand'].append(torch.rand(r, n_coarse))
        if raw_noise_std > 0:
            parts['sigma_noise_coarse'].append(noise(r * n_coarse, r, n_coarse))
        if has_fine:
            if perturb:
                parts['u_rand'].append(torch.rand(r, n_fine))
            if raw_noise_std > 0:
                parts['sigma_noise_fine'].append(noise(r * (n_coarse + n_fine), r, n_coarse + n_fine))
    return {k: torch.cat(v, 0) for k, v in parts.items() if v}


def render(state_dict: Dict[str, torch.Tensor], batch: Dict[str, torch.Tensor], *, ndc: bool,
           n_coarse: int = 64, n_fine: int = 128, retraw: bool = False, sec_views_vis: bool = False,
           white_bkgd: bool = False, lindisp: bool = False, mode: str = 'fp32', has_fine: bool = True,
           chunk: int = 4096, netchunk: int = 16384,
           forced_z_fine: Optional[torch.Tensor] = None,
           train_randoms: Optional[Dict[str, torch.Tensor]] = None) -> Dict[str, torch.Tensor]:
    """The render of a ray batch: same output keys/shapes as `VipNeRF.forward`.
    Reference: forward :34-41, batchify_rays :47-72, render_rays :74-171, run_network :264-293,
    batchify :295-329.  `forced_z_fine` (teacher forcing) replaces the fine sample positions.
    `train_randoms` (see draw_training_randoms) = train mode: stratified jitter, random cdf positions and density
    noise from the given draws; retraw and sec_views_vis are then forced on like `forward` :40 does.  The result is
    differentiable w.r.t. `state_dict` tensors that require grad (torch autograd = the gradient oracle)."""
    n = batch['rays_o'].shape[0]
    if train_randoms is not None:
        retraw, sec_views_vis = True, True
    outs = []
    for i in range(0, n, chunk):
        sub = {k: (v[i:i + chunk] if isinstance(v, torch.Tensor) and v.shape[0] == n else v)
               for k, v in batch.items()}
        fz = forced_z_fine[i:i + chunk] if forced_z_fine is not None else None
        tr = {k: v[i:i + chunk] for k, v in train_randoms.items()} if train_randoms is not None else {}
        outs.append(_render_chunk(state_dict, sub, ndc, n_coarse, n_fine, retraw, sec_views_vis, white_bkgd,
                                  lindisp, mode, has_fine, netchunk, fz, tr))
    return {k: torch.cat([o[k] for o in outs], dim=0) for k in outs[0]}


def _run_mlp(params, pts, view_dirs, view_dirs2, mode, netchunk, sigma_noise=None):
    r, s = pts.shape[:2]
    pts_flat = pts.reshape(-1, 3)
    vd_flat = view_dirs[:, None, :].expand(r, s, 3).reshape(-1, 3)
    vd2_flat = view_dirs2.reshape(r * s, view_dirs2.shape[2], 3) if view_dirs2 is not None else None
    noise_flat = sigma_noise.reshape(-1, 1) if sigma_noise is not None else None
    chunks = []
    for i in range(0, pts_flat.shape[0], netchunk):
        chunks.append(mlp_forward(params, pts_flat[i:i + netchunk], vd_flat[i:i + netchunk],
                                  vd2_flat[i:i + netchunk] if vd2_flat is not None else None, mode=mode,
                                  sigma_noise=noise_flat[i:i + netchunk] if noise_flat is not None else None))
    merged = {k: torch.cat([c[k] for c in chunks], dim=0) for k in chunks[0]}
    return {k: v.reshape(r, s, *v.shape[1:]) for k, v in merged.items()}


def _render_chunk(sd, b, ndc, n_coarse, n_fine, retraw, sec_views_vis, white_bkgd, lindisp, mode, has_fine,
                  netchunk, forced_z_fine, train=None):
    train = train or {}
    rays_o, rays_d, view_dirs = b['rays_o'], b['rays_d'], b['view_dirs']
    if ndc:
        p_o, p_d, near, far = b['rays_o_ndc'], b['rays_d_ndc'], b['near_ndc'], b['far_ndc']
    else:
        p_o, p_d, near, far = rays_o, rays_d, b['near'], b['far']
    rays_o2 = b.get('rays_o2') if sec_views_vis else None
    ret = {}

    def one_pass(tag, params, z):
        pts = p_o[..., None, :] + p_d[..., None, :] * z[..., :, None]
        vd2 = other_view_dirs(z, rays_o, rays_d, rays_o2, ndc) if rays_o2 is not None else None
        raw = _run_mlp(params, pts, view_dirs, vd2, mode, netchunk, train.get(f'sigma_noise_{tag}'))
        comp = composite(raw['sigma'][..., 0], raw['rgb'], z, p_d, ndc, rays_o, rays_d, white_bkgd,
                         raw['visibility2'][..., 0] if 'visibility2' in raw else None)
        ret[f'z_vals_{tag}'] = z
        for k, v in comp.items():
            ret[f'{k}_{tag}'] = v
        if retraw:
            ret[f'raw_sigma_{tag}'] = raw['sigma']
            ret[f'raw_rgb_view_dependent_{tag}'] = raw['rgb']
            ret[f'raw_visibility_{tag}'] = raw['visibility']
            if 'visibility2' in raw:
                ret[f'raw_visibility2_{tag}'] = raw['visibility2']
            ret[f'raw_rgb_{tag}'] = raw['rgb']
        return comp['weights']

    z_c = coarse_z_vals(near, far, n_coarse, lindisp, train.get('t_rand'))
    w_c = one_pass('coarse', split_state_dict(sd, 'coarse_model'), z_c)
    if has_fine:
        z_f = forced_z_fine if forced_z_fine is not None else fine_z_vals(z_c, w_c, n_fine, train.get('u_rand'))
        one_pass('fine', split_state_dict(sd, 'fine_model'), z_f)
    if not retraw:
        for tag in ('coarse', 'fine') if has_fine else ('coarse',):
            for k in ('z_vals', 'visibility', 'weights'):
                del ret[f'{k}_{tag}']
    return ret


def psnr_u8(pred_rgb: torch.Tensor, ref_rgb: torch.Tensor) -> float:
    """PSNR between two float rgb maps after the reference's uint8 post-processing.
    Reference: DataPreprocessor01.post_process_image :1074-1078 (clip, round(x*255)) and
    qa/02_PSNR/src/PSNR02_NeRF_LLFF.py:33-39 (10 log10(255^2 / mse))."""
    a = numpy.round(numpy.clip(pred_rgb.detach().cpu().numpy(), 0, 1) * 255).astype('uint8').astype('float64')
    b = numpy.round(numpy.clip(ref_rgb.detach().cpu().numpy(), 0, 1) * 255).astype('uint8').astype('float64')
    mse = numpy.mean((a - b) ** 2)
    return float('inf') if mse == 0 else float(10 * numpy.log10(255 ** 2 / mse))
