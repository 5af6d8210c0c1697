"""Generates tests/golden/train_*.npz: one TRAINING step of the UNMODIFIED reference (model in train mode ->
LossComputer01 with the four losses of the shipped configs -> TotalLoss.backward()), imported from /root/reference.
Run in the build container only:  python oracle/make_golden_train.py
TEST INFRASTRUCTURE ONLY.

What the fixtures pin:
 * the order in which the reference consumes torch's CPU generator in train mode (oracle.draw_training_randoms
   re-draws from the same seed; the fixture stores the draws it obtained, and the outputs only match if they are
   the reference's),
 * the train-mode forward outputs (stratified jitter, random cdf samples, density noise),
 * the loss value and the gradient of every parameter tensor - stored as the full tensors for the small ones and as
   (sum, L2 norm, 256 strided samples) for the large ones, which keeps the fixtures small.
"""
from __future__ import annotations

import copy
import os
import sys

import numpy
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import ref_loader  # noqa: E402
from oracle import vipnerf_oracle as O  # noqa: E402
from oracle.make_golden import save  # noqa: E402

SEED = 77
N_RAYS = 48
N_SEC = 2
CHUNK, NETCHUNK = 32, 1000   # small on purpose: two ray chunks and several network chunks per pass
ITER_NUM = 40000             # VisibilityPriorLoss01 is weighted 0 before iteration 30000


def grad_fingerprint(g: torch.Tensor):
    flat = g.detach().reshape(-1).double()
    if flat.numel() <= 4096:
        return {'full': g.detach()}
    idx = torch.linspace(0, flat.numel() - 1, 256).long()
    return {'sum': flat.sum(), 'norm': flat.norm(), 'samples': flat[idx].float(), 'absmax': flat.abs().max()}


def main():
    torch.set_num_threads(max(1, os.cpu_count() or 1))
    get_model = ref_loader.import_reference()
    from loss_functions.LossComputer01 import LossComputer   # the reference's own loss plumbing

    sd = O.synth_state_dict(0)
    for scene in ('fern', 'dtu'):
        ndc = O.SCENES[scene]['ndc']
        cfg = copy.deepcopy(ref_loader.reference_configs(ndc))
        cfg['model']['chunk'], cfg['model']['netchunk'] = CHUNK, NETCHUNK
        model = get_model(cfg, None)
        model.load_state_dict(sd)
        model.train()
        loss_computer = LossComputer(cfg)
        rays = O.make_rays(scene, N_RAYS, seed=9, n_sec_views=N_SEC)
        sup = O.make_supervision(scene, N_RAYS, N_SEC, ITER_NUM)
        batch = {**rays, **sup}

        torch.manual_seed(SEED)
        out = model(dict(batch))                       # train mode: retraw + sec_views_vis forced on
        losses = loss_computer.compute_losses(dict(batch), out)
        total = losses['TotalLoss']
        total.backward()

        torch.manual_seed(SEED)
        draws = O.draw_training_randoms(N_RAYS, 64, 128, CHUNK, NETCHUNK, cfg['model']['perturb'] > 0,
                                        cfg['model']['raw_noise_std'])
        arrays = {f'in.{k}': v for k, v in rays.items()}
        arrays.update({f'sup.{k}': (v if isinstance(v, torch.Tensor) else numpy.asarray(v)) for k, v in sup.items()})
        arrays.update({f'draw.{k}': v for k, v in draws.items()})
        keep = ('rgb', 'depth', 'acc', 'visibility2', 'z_vals', 'raw_sigma', 'raw_visibility', 'visibility',
                'raw_visibility2', 'raw_rgb')
        arrays.update({f'out.{k}': v for k, v in out.items() if k.rsplit('_', 1)[0] in keep})
        arrays['loss.total'] = total.detach()
        for name in ('MSE01', 'VisibilityLoss01', 'VisibilityPriorLoss01', 'SparseDepthMSE01'):
            arrays[f'loss.{name}'] = torch.as_tensor(losses[name]['loss_value']).detach()
        for name, p in model.named_parameters():
            for k, v in grad_fingerprint(p.grad).items():
                arrays[f'grad.{name}.{k}'] = v
        save(f'train_{scene}.npz', **arrays)
        print(scene, 'loss', float(total), {k: float(torch.as_tensor(v['loss_value'])) for k, v in losses.items() if k != 'TotalLoss'})


if __name__ == '__main__':
    main()
