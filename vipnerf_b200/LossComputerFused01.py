"""Drop-in for the reference's loss computer with the four ViP-NeRF losses fused into the CUDA training step
("loss fusion").

Reference interface (src/loss_functions/LossComputer01.py:12-69): `LossComputer(configs)`,
`compute_losses(input_dict, output_dict, return_loss_maps=False) -> {loss_name: {'loss_value': tensor}, ..., 'TotalLoss': tensor}`,
weights from `configs['losses'][i]['weight']` or `['iter_weights']` (:53-69); the trainer calls
`TotalLoss.backward()` (src/Trainer01.py:93-96).  Losses covered: MSE01 (MSE01.py:25-67), VisibilityLoss01
(VisibilityLoss01.py:26-74, mutual detach :57-58), VisibilityPriorLoss01 (VisibilityPriorLoss01.py:25-89),
SparseDepthMSE01 (SparseDepthMSE01.py:26-71).

When `output_dict` comes from `VipNeRFFused` in train mode (a `training.TrainOutputs`), the loss VALUES are computed by
`vipnerf_fused_losses` from the forward outputs, and `TotalLoss.backward()` makes `vipnerf_train_backward_fused` form
dTotalLoss/d(output) inside the compositing-backward kernel - the [R,S] gradient tensors of `raw_visibility` /
`visibility` that torch autograd would write and the kernel would re-read never exist.  For any other output dict
(validation with the model in eval mode, the reference model, return_loss_maps=True) the same four formulas are
evaluated with torch ops, like the reference.  Unknown loss names raise, as in the reference (:31).
"""
from __future__ import annotations

import ctypes
from typing import Dict, Optional

import torch

from . import _lib, renderpath, training

KNOWN = ('MSE01', 'VisibilityLoss01', 'VisibilityPriorLoss01', 'SparseDepthMSE01')


class LossComputer:
    def __init__(self, configs: dict):
        self.configs = configs
        self.losses: Dict[str, dict] = {}
        for loss_configs in configs['losses']:
            name = loss_configs['name']
            if name not in KNOWN:
                raise RuntimeError(f'Unknown Loss Function: {name}')
            self.losses[name] = loss_configs
        self.coarse_mlp_needed = 'coarse_mlp' in configs['model']
        self.fine_mlp_needed = 'fine_mlp' in configs['model']

    @staticmethod
    def get_loss_weight(loss_configs: dict, iter_num):
        """LossComputer01.get_loss_weight (:53-69)."""
        weight = None
        if 'weight' in loss_configs:
            weight = loss_configs['weight']
        elif 'iter_weights' in loss_configs:
            for key in sorted((int(k) for k in loss_configs['iter_weights']), reverse=True):
                if iter_num >= key:
                    weight = loss_configs['iter_weights'][str(key)]
                    break
        if weight is None:
            raise RuntimeError(f"loss_weight is None for {loss_configs['name']} at iter {iter_num}")
        return weight

    # ------------------------------------------------------------------ the reference's entry point
    def compute_losses(self, input_dict: dict, output_dict: dict, return_loss_maps: bool = False):
        if 'common_data' in input_dict.keys():   # LossComputer01.py:34-38
            for key in input_dict['common_data'].keys():
                if isinstance(input_dict['common_data'][key], torch.Tensor) and input_dict['common_data'][key].dim() > 3:
                    input_dict['common_data'][key] = input_dict['common_data'][key][0]
        fused = getattr(output_dict, 'fused', None)
        if fused is not None and not return_loss_maps and isinstance(output_dict, training.TrainOutputs):
            return self._compute_fused(input_dict, output_dict, fused)
        return self._compute_torch(input_dict, output_dict, return_loss_maps)

    def _active(self, input_dict: dict, output_dict: dict) -> Dict[str, float]:
        """name -> weight of the losses that contribute for this batch (the reference skips VisibilityPriorLoss01 when
        the model returned no raw_visibility2, VisibilityPriorLoss01.py:29-31)."""
        iter_num = input_dict['iter_num']
        active = {}
        for name, cfg in self.losses.items():
            if name == 'VisibilityPriorLoss01' and (
                    (self.coarse_mlp_needed and 'raw_visibility2_coarse' not in output_dict)
                    or (self.fine_mlp_needed and 'raw_visibility2_fine' not in output_dict)):
                continue
            active[name] = float(self.get_loss_weight(cfg, iter_num))
        return active

    # ------------------------------------------------------------------ fused path
    def _compute_fused(self, input_dict: dict, output_dict: 'training.TrainOutputs', fused: dict):
        lib = _lib.load()
        active = self._active(input_dict, output_dict)
        rgb = output_dict['rgb_coarse']
        device, R = rgb.device, rgb.shape[0]
        keep = []

        def f32(t, name):
            t = renderpath._f32c(t.to(device), name)
            keep.append(t)
            return t

        def mask(t):
            t = t.to(device).to(torch.bool).contiguous()
            keep.append(t)
            return t

        spec = _lib.LossSpec()
        spec.w_mse = active.get('MSE01', 0.0)
        spec.w_visibility = active.get('VisibilityLoss01', 0.0)
        spec.w_prior = active.get('VisibilityPriorLoss01', 0.0)
        spec.w_sparse_depth = active.get('SparseDepthMSE01', 0.0)
        if 'MSE01' in active or 'VisibilityPriorLoss01' in active:
            spec.mask_nerf = mask(input_dict['indices_mask_nerf']).data_ptr()
        if 'MSE01' in active:
            spec.target_rgb = f32(input_dict['target_rgb'], 'target_rgb').data_ptr()
        if 'VisibilityPriorLoss01' in active:
            prior = input_dict.get('visibility_prior_masks', input_dict.get('visibility_prior_weights'))
            if prior is not None:     # else: ones (VisibilityPriorLoss01.py:38-41)
                spec.prior = f32(prior, 'visibility_prior_masks').data_ptr()
        if 'SparseDepthMSE01' in active and 'indices_mask_sparse_depth' in input_dict:
            spec.mask_sparse_depth = mask(input_dict['indices_mask_sparse_depth']).data_ptr()
            spec.sparse_depth = f32(input_dict['sparse_depth_values'][:, 0], 'sparse_depth_values').data_ptr()
        fwd = _lib.Out()
        for name, t in output_dict.items():
            key, tag = name.rsplit('_', 1)
            if key in _lib.PASS_FIELDS and isinstance(t, torch.Tensor):
                setattr(getattr(fwd, tag), key, t.data_ptr())
        losses_dev = torch.empty(8, dtype=torch.float32, device=device)
        ws = torch.empty(max(1, (R + 3) // 4) * 16, dtype=torch.float32, device=device)
        cfg = _lib.make_cfg(**fused['cfg_kwargs'])
        with torch.cuda.device(device):
            _lib.check(lib.vipnerf_fused_losses(ctypes.byref(cfg), R, ctypes.byref(fwd), ctypes.byref(spec),
                                                losses_dev.data_ptr(), ws.data_ptr(), ws.numel() * 4,
                                                renderpath._stream(device)), 'vipnerf_fused_losses')
        spec.losses_dev = losses_dev.data_ptr()
        keep.append(losses_dev)
        state = fused['state']
        state.loss_spec, state.keepalive = spec, keep
        result = {}
        for i, name in enumerate(KNOWN):
            if name in active:
                result[name] = {'loss_value': losses_dev[i]}
        result['TotalLoss'] = training._FusedTotal.apply(fused['token'], losses_dev)
        return result

    # ------------------------------------------------------------------ torch path (validation, foreign outputs, loss maps)
    def _compute_torch(self, input_dict: dict, output_dict: dict, return_loss_maps: bool):
        active = self._active(input_dict, output_dict)
        tags = (['coarse'] if self.coarse_mlp_needed else []) + (['fine'] if self.fine_mlp_needed else [])
        values, maps = {}, {}

        def masked_mean(x, m):
            x = x[m]
            return torch.mean(x) if x.numel() > 0 else 0

        if 'MSE01' in active:
            m = input_dict['indices_mask_nerf']
            total = 0
            for t in tags:
                mse = torch.mean(torch.square(output_dict[f'rgb_{t}'] - input_dict['target_rgb']), dim=1)
                total = total + masked_mean(mse, m)
                maps[f'MSE01_{t}'] = mse[m]
            values['MSE01'] = total
        if 'VisibilityLoss01' in active:
            total = 0
            for t in tags:
                pred, tgt = output_dict[f'raw_visibility_{t}'][..., 0], output_dict[f'visibility_{t}']
                l1 = torch.mean(torch.abs(pred - tgt.detach()), dim=1)
                l2 = torch.mean(torch.abs(pred.detach() - tgt), dim=1)
                total = total + torch.mean(l1) + torch.mean(l2)
                maps[f'VisibilityLoss01_{t}'] = l1 + l2
            values['VisibilityLoss01'] = total
        if 'VisibilityPriorLoss01' in active:
            m = input_dict['indices_mask_nerf']
            prior = input_dict.get('visibility_prior_masks', input_dict.get('visibility_prior_weights'))
            total = 0
            for t in tags:
                vis2 = output_dict[f'visibility2_{t}']
                pw = prior if prior is not None else torch.ones_like(vis2)
                per_ray = torch.sum(pw * (1 - vis2), dim=1)
                total = total + masked_mean(per_ray, m)
                maps[f'VisibilityPriorLoss01_{t}'] = per_ray[m]
            values['VisibilityPriorLoss01'] = total
        if 'SparseDepthMSE01' in active:
            if 'indices_mask_sparse_depth' not in input_dict:
                values['SparseDepthMSE01'] = torch.zeros((), device=input_dict['rays_o'].device)
            else:
                t = 'fine' if self.fine_mlp_needed else 'coarse'
                err = torch.square(output_dict[f'depth_{t}'] - input_dict['sparse_depth_values'][:, 0])
                values['SparseDepthMSE01'] = masked_mean(err, input_dict['indices_mask_sparse_depth'])
        result, total = {}, 0
        for name, v in values.items():
            entry = {'loss_value': v}
            if return_loss_maps:
                entry['loss_maps'] = {k: m for k, m in maps.items() if k.startswith(name)}
            result[name] = entry
            total = total + active[name] * v
        result['TotalLoss'] = total
        return result
