"""Drop-in model plugin: the reference's `VipNeRF` module with `render_rays` executed by the B200-native
CUDA library instead of the PyTorch graph.

Plugin contract (reference src/models/ModelFactory.py:10-22): file `<Name>NN.py` holding class `<Name>`,
constructed as `<Name>(configs, model_configs)`; here `configs['model']['name'] = 'VipNeRFFused01'`.
Same `forward(input_batch, retraw=False, sec_views_vis=False) -> dict` signature, output keys and shapes
as `VipNeRF.forward` (src/models/VipNeRF01.py:34-41, :128-133, :161-170, :366-383), same `state_dict` keys
(`coarse_model.pts_linears.0.weight`, ... ) so reference checkpoints load unchanged
(src/Tester01.py:45-49), usable under torch.nn.DataParallel (src/Tester01.py:42).

`model.eval()`: the fused tensor-core render (configs['model']['precision'], default bf16).  `model.train()`: the
training step of SURVEY.md section 8 row f1 - train-mode forward (stratified jitter, random cdf samples, density noise
drawn from torch's CPU generator in the reference's order; retraw and sec_views_vis forced on, VipNeRF01.py:40) and its
backward through vipnerf_b200.training, so that `loss.backward()` fills `.grad` of the same parameters the reference's
optimizer steps (Trainer01.py:93-102, :519) - in fp32 like the reference's training arithmetic by default, or with every
256-wide product on the tensor cores (configs['model']['train_precision'] = 'tf32').
There is no CPU fallback: inputs must be CUDA tensors and the shared library must be built.
"""
from __future__ import annotations

import threading
from typing import Dict, Optional

import torch

from . import renderpath, training


class RadianceMLPParams(torch.nn.Module):
    """Parameter container with the reference MLP's module names and shapes (VipNeRF01.py:472-491).  It owns the
    fp32 master weights only; evaluation happens in the CUDA library on a packed copy."""

    def __init__(self, mlp_configs: dict):
        super().__init__()
        W = mlp_configs['netwidth']
        D = mlp_configs['netdepth']
        pts_dim = 3 + 6 * mlp_configs['points_positional_encoding_degree']
        views_dim = 3 + 6 * mlp_configs['views_positional_encoding_degree']
        if not (mlp_configs['use_view_dirs'] and mlp_configs['view_dependent_rgb'] and mlp_configs['predict_visibility']):
            raise NotImplementedError('VipNeRFFused is built for ViP-NeRF MLPs with use_view_dirs, view_dependent_rgb '
                                      'and predict_visibility all enabled (every shipped config)')
        skips = [4]
        self.pts_linears = torch.nn.ModuleList(
            [torch.nn.Linear(pts_dim, W)]
            + [torch.nn.Linear(W + pts_dim if i in skips else W, W) for i in range(D - 1)])
        self.views_linears = torch.nn.ModuleList([torch.nn.Linear(views_dim + W, W // 2)])
        self.pts_output_linear = torch.nn.Linear(W, 1)
        self.feature_linear = torch.nn.Linear(W, W)
        self.views_output_linear = torch.nn.Linear(W // 2, 4)
        self.mlp_configs = mlp_configs

    def forward(self, *args, **kwargs):
        raise RuntimeError('RadianceMLPParams holds parameters only; call VipNeRFFused.forward')

    def named_tensors(self) -> Dict[str, torch.Tensor]:
        """The 24 tensors under the reference's state_dict names, read BY ATTRIBUTE: inside an nn.DataParallel replica
        (Tester01.py:42, Trainer01.py:517 with a device list) `_parameters` is empty - `parameters()` and `state_dict()`
        return nothing there - while `linear.weight` is the replica's broadcast copy (a non-leaf tensor that carries the
        autograd edge back to the master parameter)."""
        out: Dict[str, torch.Tensor] = {}
        for i, layer in enumerate(self.pts_linears):
            out[f'pts_linears.{i}.weight'], out[f'pts_linears.{i}.bias'] = layer.weight, layer.bias
        for name, layer in (('views_linears.0', self.views_linears[0]), ('pts_output_linear', self.pts_output_linear),
                            ('feature_linear', self.feature_linear), ('views_output_linear', self.views_output_linear)):
            out[f'{name}.weight'], out[f'{name}.bias'] = layer.weight, layer.bias
        return out


class _PackCache:
    """Packed weight images per (MLP, precision, device).  Never copied or pickled with the module: copy.deepcopy(model)
    and torch.save(model) get a fresh, empty cache (a lock and CUDA events are not copyable)."""

    def __init__(self):
        self.lock = threading.Lock()
        self.entries: Dict[tuple, tuple] = {}

    def __deepcopy__(self, memo):
        return _PackCache()

    def __reduce__(self):
        return (_PackCache, ())


class VipNeRFFused(torch.nn.Module):
    def __init__(self, configs: dict, model_configs: Optional[dict] = None):
        super().__init__()
        self.configs = configs
        self.model_configs = model_configs
        model_cfg = configs['model']
        self.ndc = configs['data_loader']['ndc']
        self.coarse_mlp_needed = 'coarse_mlp' in model_cfg
        self.fine_mlp_needed = 'fine_mlp' in model_cfg
        if not self.coarse_mlp_needed:
            raise NotImplementedError('a coarse MLP is required')
        self.precision = model_cfg.get('precision', 'bf16')   # 'fp32' | 'bf16' | 'fp16' | 'bf16x3'
        if self.precision not in ('fp32', 'bf16', 'fp16', 'bf16x3'):
            raise ValueError(f"configs['model']['precision'] = {self.precision!r}")
        # training arithmetic: 'fp32' (like the reference), 'tf32' = the 256-wide products of the step on the tensor cores,
        # 'fp16' = the same products with fp16 saved activations / chain gradients (half the HBM traffic of the step)
        if model_cfg.get('train_precision', 'fp32') not in ('fp32', 'tf32', 'fp16'):
            raise ValueError(f"configs['model']['train_precision'] = {model_cfg['train_precision']!r}")
        self.coarse_model = RadianceMLPParams(model_cfg['coarse_mlp'])
        self.fine_model = RadianceMLPParams(model_cfg['fine_mlp']) if self.fine_mlp_needed else None
        for name in ('coarse_mlp', 'fine_mlp'):
            if name in model_cfg:
                m = model_cfg[name]
                shape = (m['netdepth'], m['netwidth'], m['points_positional_encoding_degree'],
                         m['views_positional_encoding_degree'])
                if shape != (8, 256, 10, 4):
                    raise NotImplementedError(f'{name} shape {shape}: kernels are built for (8, 256, 10, 4)')
        # Shared BY REFERENCE between nn.DataParallel replicas (replicate() copies __dict__ shallowly): entries are keyed
        # per device and validated against the tensors the calling replica actually holds, so sharing is harmless.
        self._pack_cache = _PackCache()

    # ------------------------------------------------------------------ packed-weight cache
    def invalidate_packed(self) -> None:
        """Drops the packed weight images.  Needed only after writes that bypass autograd's version counter
        (`p.data.copy_()`, `p.data.mul_()` ...); optimizer steps, load_state_dict and .to() are detected."""
        with self._pack_cache.lock:
            self._pack_cache.entries.clear()

    def weights_version(self) -> tuple:
        """(storage address, in-place version) of every parameter tensor: changes whenever an optimizer step,
        load_state_dict or .to() touches a weight.  hostio.GraphedRender re-captures its CUDA graph when it changes (the
        packed weight images a captured launch points to belong to one version)."""
        models = [self.coarse_model] + ([self.fine_model] if self.fine_mlp_needed else [])
        return tuple((t.data_ptr(), t._version) for m in models for t in m.named_tensors().values())

    def _packed_weights(self, which: str, precision: str, device) -> torch.Tensor:
        mlp = self.coarse_model if which == 'coarse' else self.fine_model
        tensors = mlp.named_tensors()
        # (storage address, in-place version) of every tensor: cheap to read, changes whenever an optimizer step,
        # load_state_dict or .to() touches a weight - and on every forward of a DataParallel replica, whose weights are
        # fresh broadcast copies
        version = tuple([(t.data_ptr(), t._version) for t in tensors.values()])
        key = (which, precision, device.index)
        stream = torch.cuda.current_stream(device)
        if getattr(mlp, '_is_replica', False):
            # a replica's tensors live for one forward only; a later broadcast may reuse their addresses with other
            # values, so (address, version) identifies nothing there: pack per call (2.6 MB, one small kernel)
            return renderpath.pack_mlp({k: v.detach() for k, v in tensors.items()}, precision)
        with self._pack_cache.lock:
            hit = self._pack_cache.entries.get(key)
            if hit is not None and hit[0] == version:
                # packed on another stream: order this stream after the pack kernel (not while capturing a CUDA graph:
                # hostio.GraphedRender synchronises the device before it captures)
                if hit[2] != stream.cuda_stream and not torch.cuda.is_current_stream_capturing():
                    stream.wait_event(hit[3])
                return hit[1]
            packed = renderpath.pack_mlp({k: v.detach() for k, v in tensors.items()}, precision)
            done = torch.cuda.Event()
            done.record(stream)
            self._pack_cache.entries[key] = (version, packed, stream.cuda_stream, done)
            return packed

    # ------------------------------------------------------------------ the reference's forward contract
    def forward(self, input_batch: dict, retraw: bool = False, sec_views_vis: bool = False, out: Optional[dict] = None):
        """`out` (optional, eval mode; not part of the reference signature): name -> preallocated fp32 CUDA tensor the
        kernel writes that output into, e.g. views of another GPU's memory (sharding.PeerGather)."""
        if 'common_data' in input_batch.keys():   # VipNeRF01.py:35-39
            for key in input_batch['common_data'].keys():
                if isinstance(input_batch['common_data'][key], torch.Tensor):
                    input_batch['common_data'][key] = input_batch['common_data'][key][0]
        if self.training:
            return self.render_train(input_batch)
        return self.render(input_batch, retraw=retraw, sec_views_vis=sec_views_vis, out=out)

    def _ray_batch(self, input_dict: dict, sec_views_vis: bool):
        rays_o = input_dict['rays_o']
        if not isinstance(rays_o, torch.Tensor) or not rays_o.is_cuda:
            raise RuntimeError('VipNeRFFused needs CUDA tensors (move the batch with CommonUtils.move_to_device); '
                               'there is no CPU fallback')
        batch = {k: input_dict[k] for k in ('rays_o', 'rays_d', 'view_dirs', 'near', 'far', 'rays_o_ndc',
                                            'rays_d_ndc', 'near_ndc', 'far_ndc') if k in input_dict}
        n_sec_views = 0
        if sec_views_vis:   # VipNeRF01.py:84-98
            if 'rays_o2' in input_dict:
                rays_o2 = input_dict['rays_o2']
            else:
                poses = input_dict['common_data']['poses']
                image_id = input_dict['pixel_id'][:, 0].long()
                others = [poses[i + (i >= image_id).long()][:, :3, 3] for i in range(input_dict['num_frames'] - 1)]
                rays_o2 = torch.stack(others, dim=1)
            batch['rays_o2'] = rays_o2
            n_sec_views = rays_o2.shape[1]
        return batch, n_sec_views

    def render_train(self, input_dict: dict):
        """Train-mode forward (VipNeRF01.py:40: retraw and sec_views_vis forced on), differentiable w.r.t. the
        parameters; random numbers are drawn here, from torch's CPU generator, where the reference draws them."""
        model_cfg = self.configs['model']
        batch, n_sec_views = self._ray_batch(input_dict, True)
        device = batch['rays_o'].device
        n_coarse = model_cfg['coarse_mlp']['num_samples']
        n_fine = model_cfg['fine_mlp']['num_samples'] if self.fine_mlp_needed else 0
        n_rays = batch['rays_o'].shape[0]
        perturb, noise_std = model_cfg['perturb'] > 0, float(model_cfg['raw_noise_std'])
        if model_cfg.get('rng', 'reference') == 'device':
            # configs['model']['rng'] = 'device': the same distributions drawn by the device generator - no host work
            # and no upload, but not the reference's random stream (a seeded run differs from the reference's)
            draws = {}
            if perturb:
                draws['t_rand'] = torch.rand(n_rays, n_coarse, device=device)
                if self.fine_mlp_needed:
                    draws['u_rand'] = torch.rand(n_rays, n_fine, device=device)
            if noise_std > 0:
                draws['sigma_noise_coarse'] = torch.randn(n_rays, n_coarse, device=device) * noise_std
                if self.fine_mlp_needed:
                    draws['sigma_noise_fine'] = torch.randn(n_rays, n_coarse + n_fine, device=device) * noise_std
        else:
            draws = training.draw_training_randoms(n_rays, n_coarse, n_fine, model_cfg['chunk'], model_cfg['netchunk'],
                                                   perturb, noise_std, self.fine_mlp_needed)
        batch.update({k: v.to(device, non_blocking=True) for k, v in draws.items()})
        return training.render_rays_train(
            batch, self.coarse_model.named_tensors(), self.fine_model.named_tensors() if self.fine_mlp_needed else None,
            ndc=self.ndc, n_coarse=n_coarse, n_fine=n_fine, n_sec_views=n_sec_views,
            white_bkgd=model_cfg['white_bkgd'], lindisp=model_cfg['lindisp'],
            train_precision=model_cfg.get('train_precision', 'fp32'))

    def render(self, input_dict: dict, retraw: bool, sec_views_vis: bool, out: Optional[dict] = None):
        batch, n_sec_views = self._ray_batch(input_dict, sec_views_vis)
        device = batch['rays_o'].device
        model_cfg = self.configs['model']
        # visibility2 runs on the tensor path too (one K=32 MMA step per secondary view); more than 8 views per
        # tile only fit the fp32 kernels
        precision = 'fp32' if n_sec_views > 8 else self.precision
        packed_c = self._packed_weights('coarse', precision, device)
        packed_f = self._packed_weights('fine', precision, device) if self.fine_mlp_needed else None
        if not self.fine_mlp_needed and not retraw:
            raise KeyError('z_vals_fine')   # what the reference does for a coarse-only model, VipNeRF01.py:170
        out = renderpath.render_rays(
            batch, packed_c, packed_f, ndc=self.ndc, precision=precision,
            n_coarse=model_cfg['coarse_mlp']['num_samples'],
            n_fine=model_cfg['fine_mlp']['num_samples'] if self.fine_mlp_needed else 0,
            retraw=retraw, n_sec_views=n_sec_views, white_bkgd=model_cfg['white_bkgd'], lindisp=model_cfg['lindisp'],
            out_tensors=out)
        return out
