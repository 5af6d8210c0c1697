"""A small pass over the round-2 kernels for `compute-sanitizer --tool memcheck python tools/sanitize_step.py`:
eval renders (bf16 / fp16 / bf16x3, ragged ray count, out= tensors), a tensor-core training iteration with the fused
losses in the tf32 and in the fp16 mode (TMA-staged chain epilogues, bit masks, column sums inside the parameter-gradient
products, narrow and grouped products), the training-batch gather."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import numpy  # noqa: E402
import torch  # noqa: E402

import bench  # noqa: E402
from oracle import vipnerf_oracle as O  # noqa: E402
from oracle.make_golden_train_batch import synthetic_tables  # noqa: E402
from vipnerf_b200.LossComputerFused01 import LossComputer  # noqa: E402
from vipnerf_b200.ModelFactory import get_model  # noqa: E402
from vipnerf_b200.TrainBatchFused01 import TrainBatchLoaderFused  # noqa: E402

dev = torch.device('cuda:0')
sd = O.synth_state_dict(0)
rays = {k: v.to(dev) for k, v in O.make_rays('fern', 301, seed=2, n_sec_views=1).items()}
for precision in ('bf16', 'fp16', 'bf16x3'):
    model = get_model(bench.model_configs(precision), None)
    model.load_state_dict(sd)
    model = model.to(dev).eval()
    with torch.no_grad():
        out = model(dict(rays), retraw=True, sec_views_vis=True)
        mine = {'rgb_fine': torch.empty(301, 3, device=dev)}
        model(dict(rays), out=mine)
    torch.cuda.synchronize()
    print('eval', precision, float(out['rgb_fine'].mean()), float(mine['rgb_fine'].mean()))
batch = dict(rays)
batch.update({k: (v.to(dev) if isinstance(v, torch.Tensor) else v) for k, v in O.make_supervision('fern', 301, 1).items()})
for train_precision in ('tf32', 'fp16'):     # fp16: fp16 arrays, bit masks, heads in the product epilogues, grouped products
    cfg = bench.model_configs('bf16')
    cfg['model'].update(rng='device', train_precision=train_precision)
    cfg['losses'] = [{'name': 'MSE01', 'weight': 1}, {'name': 'VisibilityLoss01', 'weight': 0.1},
                     {'name': 'VisibilityPriorLoss01', 'weight': 0.001}, {'name': 'SparseDepthMSE01', 'weight': 0.1}]
    model = get_model(cfg, None)
    model.load_state_dict(sd)
    model = model.to(dev).train()
    losses = LossComputer(cfg).compute_losses(batch, model(dict(batch)))
    losses['TotalLoss'].backward()
    torch.cuda.synchronize()
    print('train', train_precision, float(losses['TotalLoss']), float(model.coarse_model.pts_linears[0].bias.grad.abs().sum()))
tables = synthetic_tables(3, True, True, n_frames=2, h=20, w=30)
loader = TrainBatchLoaderFused(tables, device=dev, ndc=True, num_rays=333, num_rays_sparse_depth=77, prior_masks=True)
numpy.random.seed(0)
for it in range(3):
    b = loader.load_cached_next_batch(it, None)
torch.cuda.synchronize()
print('gather', float(b['rays_o'].sum()))
print('sanitize_step OK')
