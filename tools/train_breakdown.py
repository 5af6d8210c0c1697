"""Where a training step's time goes: wall-clock (with a device sync after each phase) of the random draws, the
train-mode forward, the losses, the backward and the optimizer step.   python tools/train_breakdown.py [--rays 4096]"""
import argparse
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import torch  # noqa: E402

import bench  # noqa: E402
from oracle import vipnerf_oracle as O  # noqa: E402
from vipnerf_b200 import training  # noqa: E402
from vipnerf_b200.ModelFactory import get_model  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument('--rays', type=int, default=4096)
ap.add_argument('--steps', type=int, default=4)
ap.add_argument('--rng', default='reference')
ap.add_argument('--train-precision', default='fp32')
args = ap.parse_args()

cfg = bench.model_configs('bf16', ndc=True)
cfg['model']['rng'] = args.rng
cfg['model']['train_precision'] = args.train_precision
model = get_model(cfg, None)
model.load_state_dict(O.synth_state_dict(0))
model = model.cuda().train()
opt = torch.optim.Adam(model.parameters(), lr=5e-4)
rays = {k: v.cuda() for k, v in O.make_rays('re10k', args.rays, seed=2, n_sec_views=1).items()}
target = O.make_supervision('re10k', args.rays, 1)['target_rgb'].cuda()


def timed(fn):
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    r = fn()
    torch.cuda.synchronize()
    return r, (time.perf_counter() - t0) * 1e3


for step in range(args.steps):
    _, t_draw = timed(lambda: training.draw_training_randoms(args.rays, 64, 128, 4096, 16384, True, 1.0, True))
    opt.zero_grad(set_to_none=True)
    out, t_fwd = timed(lambda: model(dict(rays)))
    loss, t_loss = timed(lambda: torch.mean(torch.square(out['rgb_fine'] - target)) + torch.mean(torch.square(out['rgb_coarse'] - target))
                         + 0.1 * torch.mean(torch.abs(out['raw_visibility_fine'][..., 0] - out['visibility_fine'].detach()))
                         + 0.001 * out['visibility2_fine'].mean() + 0.1 * out['depth_fine'].mean())
    _, t_bwd = timed(lambda: loss.backward())
    _, t_opt = timed(lambda: opt.step())
    print(f'step {step}: draws alone {t_draw:.1f} ms | forward (incl. its draws) {t_fwd:.1f} | losses {t_loss:.1f} | '
          f'backward {t_bwd:.1f} | adam {t_opt:.1f} | sum {t_fwd + t_loss + t_bwd + t_opt:.1f} ms')
