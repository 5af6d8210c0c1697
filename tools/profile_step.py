"""Runs a few steps of the hot path (one fused kernel per step) for ncu captures.
    ncu --set full --clock-control none --import-source on -k regex:k_render_tc -s 3 -c 1 -o gpurun_out/prof \
        python tools/profile_step.py [--precision bf16] [--rays 4096] [--steps 5]
"""
import argparse
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import torch  # noqa: E402

from oracle import vipnerf_oracle as O  # noqa: E402
from vipnerf_b200 import renderpath  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument('--precision', default='bf16')
ap.add_argument('--rays', type=int, default=4096)
ap.add_argument('--steps', type=int, default=5)
ap.add_argument('--scene', default='fern')
args = ap.parse_args()

ndc = O.SCENES[args.scene]['ndc']
sd = {k: v.cuda() for k, v in O.synth_state_dict(0).items()}
batch = {k: v.cuda() for k, v in O.make_rays(args.scene, args.rays, seed=2).items()}
pc = renderpath.pack_mlp(O.split_state_dict(sd, 'coarse_model'), args.precision)
pf = renderpath.pack_mlp(O.split_state_dict(sd, 'fine_model'), args.precision)
keys = renderpath.pass_keys(ndc, False, 0)
for _ in range(args.steps):
    out = renderpath.render_rays(batch, pc, pf, ndc=ndc, precision=args.precision, keys=keys)
torch.cuda.synchronize()
print('ok', float(out['rgb_fine'].mean()))
