#!/usr/bin/env python
"""Error of candidate tensor-core arithmetics for the render path, measured on the CPU with the oracle's emulation
modes (oracle/vipnerf_oracle.py::linear).  Answers VERDICT r01 item 2: which arithmetic meets "<= 1e-4 relative on
RGB / depth / visibility" and at how many MMAs per product.

    python tools/precision_study.py [--rays 1024] > profiles/r02_precision_study.md

Columns: MMAs per 256-wide product (the tensor-pipe cost relative to bf16), then per output key max / p99 / median of
|mode - fp32| / max|fp32|, end to end (fine samples re-drawn from the mode's own coarse weights) and teacher-forced
(the fine pass fed the fp32 run's z_vals_fine - free of sample_pdf's discontinuity).
"""
import argparse
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import vipnerf_oracle as O   # noqa: E402

MODES = [('bf16', 1), ('fp16', 1), ('tf32', 2), ('bf16_a2', 2), ('fp16_a2', 2), ('bf16x3', 3)]
KEYS = ('rgb_fine', 'depth_fine', 'depth_ndc_fine', 'acc_fine', 'visibility2_fine', 'rgb_coarse', 'depth_coarse')


def stats(a, b):
    d = ((a.double() - b.double()).abs() / b.abs().max().clamp_min(1e-30)).flatten()
    return d.max().item(), torch.quantile(d, 0.99).item(), d.median().item()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--rays', type=int, default=1024)
    ap.add_argument('--scene', default='fern')
    args = ap.parse_args()
    torch.set_num_threads(os.cpu_count() or 1)
    sd = O.synth_state_dict(0)
    ndc = O.SCENES[args.scene]['ndc']
    batch = O.make_rays(args.scene, args.rays, seed=2, n_sec_views=2)
    with torch.no_grad():
        ref = O.render(sd, batch, ndc=ndc, retraw=True, sec_views_vis=True)
        print(f'# Precision study ({args.scene}, {args.rays} rays, 2 secondary views, synthetic weights with density gain 300)\n')
        print('error = |mode - fp32 oracle| / max|fp32 oracle|, shown as max / p99 / median\n')
        for forced in (False, True):
            print(f'## {"teacher-forced fine pass (fp32 z_vals_fine)" if forced else "end to end"}\n')
            print('| mode | MMAs | ' + ' | '.join(k for k in KEYS if not (forced and k.endswith('coarse'))) + ' |')
            print('|---|---|' + '---|' * len([k for k in KEYS if not (forced and k.endswith('coarse'))]))
            for mode, cost in MODES:
                out = O.render(sd, batch, ndc=ndc, retraw=True, sec_views_vis=True, mode=mode,
                               forced_z_fine=ref['z_vals_fine'] if forced else None)
                cells = []
                for k in KEYS:
                    if forced and k.endswith('coarse'):
                        continue
                    if k not in ref:
                        cells.append('-')
                        continue
                    mx, p99, med = stats(out[k], ref[k])
                    cells.append(f'{mx:.1e} / {p99:.1e} / {med:.1e}')
                print(f'| {mode} | {cost} | ' + ' | '.join(cells) + ' |', flush=True)
            print()


if __name__ == '__main__':
    main()
