"""Training arithmetic modes against each other at the benchmarked shape: one 4096-ray iteration of BASELINE config 3
(RealEstate-10K camera, 1 secondary view, the four fused losses) from the same weights and the same device-side draws in
every `train_precision`; prints the loss values and, per parameter tensor, the gradient difference to the fp32 CUDA-core
mode (max-norm relative to the tensor's largest entry, and L2).   python tools/train_mode_compare.py [--rays 4096]"""
import argparse
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import torch  # noqa: E402

import bench  # noqa: E402
from oracle import vipnerf_oracle as O  # noqa: E402
from vipnerf_b200.LossComputerFused01 import LossComputer  # noqa: E402
from vipnerf_b200.ModelFactory import get_model  # noqa: E402


def one_step(train_precision, rays, sup, seed):
    cfg = bench.model_configs('bf16', ndc=True)
    cfg['model']['rng'] = 'device'
    cfg['model']['train_precision'] = train_precision
    cfg['losses'] = [{'name': 'MSE01', 'weight': 1}, {'name': 'VisibilityLoss01', 'weight': 0.1},
                     {'name': 'VisibilityPriorLoss01', 'iter_weights': {'0': 0, '30000': 0.001}},
                     {'name': 'SparseDepthMSE01', 'weight': 0.1}]
    model = get_model(cfg, None)
    model.load_state_dict(O.synth_state_dict(0))
    model = model.cuda().train()
    batch = dict(rays)
    batch.update(sup)
    torch.manual_seed(seed)
    out = model(batch)
    losses = LossComputer(cfg).compute_losses(batch, out)
    losses['TotalLoss'].backward()
    torch.cuda.synchronize()
    values = {k: float(v['loss_value'] if isinstance(v, dict) else v) for k, v in losses.items()}
    return values, {k: p.grad.detach().clone() for k, p in model.named_parameters()}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--rays', type=int, default=4096)
    ap.add_argument('--modes', default='tf32,fp16')
    ap.add_argument('--json', default=None)
    args = ap.parse_args()
    rays = {k: v.cuda() for k, v in O.make_rays('re10k', args.rays, seed=2, n_sec_views=1).items()}
    sup = {k: (v.cuda() if isinstance(v, torch.Tensor) else v) for k, v in O.make_supervision('re10k', args.rays, 1).items()}
    ref_values, ref_grads = one_step('fp32', rays, sup, seed=7)
    print('fp32 losses', ref_values)
    report = {'rays': args.rays, 'fp32_losses': ref_values, 'modes': {}}
    for mode in args.modes.split(','):
        values, grads = one_step(mode, rays, sup, seed=7)
        worst_max, worst_l2, worst_name = 0.0, 0.0, ''
        for k, g in ref_grads.items():
            d = (grads[k] - g).double()
            e_max = (d.abs().max() / g.abs().max().clamp_min(1e-30)).item()
            e_l2 = (d.norm() / g.double().norm().clamp_min(1e-30)).item()
            if e_max > worst_max:
                worst_max, worst_name = e_max, k
            worst_l2 = max(worst_l2, e_l2)
        finite = all(torch.isfinite(g).all().item() for g in grads.values())
        print(f'{mode}: losses {values}\n   worst gradient error vs fp32 {worst_max:.3e} ({worst_name}), worst L2 error {worst_l2:.3e}, '
              f'finite {finite}')
        report['modes'][mode] = {'losses': values, 'worst_grad_err_vs_fp32': worst_max, 'worst_tensor': worst_name,
                                 'worst_l2_err_vs_fp32': worst_l2, 'finite': finite}
    if args.json:
        with open(args.json, 'w') as f:
            json.dump(report, f, indent=1)


if __name__ == '__main__':
    main()
