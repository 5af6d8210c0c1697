"""Cycle breakdown of the fused tcgen05 kernel's warp roles (CTA 0), via vipnerf_debug_set_profile_buffer.
    python tools/tc_cycle_breakdown.py [--precision bf16] [--rays 4096]
Not a benchmark: the counters are clock64 deltas around the stages of one epilogue thread per slot and of the
MMA-issuing lane; they show where a tile's time goes."""
import argparse
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import torch  # noqa: E402

from oracle import vipnerf_oracle as O  # noqa: E402
from vipnerf_b200 import _lib, renderpath  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument('--precision', default='bf16')
ap.add_argument('--rays', type=int, default=4096)
args = ap.parse_args()

sd = {k: v.cuda() for k, v in O.synth_state_dict(0).items()}
batch = {k: v.cuda() for k, v in O.make_rays('fern', args.rays, seed=2).items()}
pc = renderpath.pack_mlp(O.split_state_dict(sd, 'coarse_model'), args.precision)
pf = renderpath.pack_mlp(O.split_state_dict(sd, 'fine_model'), args.precision)
keys = renderpath.pass_keys(True, False, 0)
for _ in range(3):
    renderpath.render_rays(batch, pc, pf, ndc=True, precision=args.precision, keys=keys)
torch.cuda.synchronize()
buf = torch.zeros(64, dtype=torch.int64, device='cuda')
lib = _lib.load()
lib.vipnerf_debug_set_profile_buffer(buf.data_ptr())
s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
s.record()
renderpath.render_rays(batch, pc, pf, ndc=True, precision=args.precision, keys=keys)
e.record()
torch.cuda.synchronize()
lib.vipnerf_debug_set_profile_buffer(None)
b = buf.cpu().tolist()
print(f'kernel {s.elapsed_time(e):.3f} ms, precision {args.precision}, rays {args.rays}')
names = ['encode', 'view_bias', 'wait_d_ready', 'layer_epilogues', 'view_epilogue+out', 'hooks', 'n_tiles', 'total']
for slot in range(2):
    q = b[slot * 16: slot * 16 + 8]
    if q[6] == 0:
        continue
    print(f'slot {slot}: tiles {q[6]}, total {q[7]} cycles = {q[7] / q[6]:.0f} / tile (tensor time per tile: 17152 cycles = 60 chunks x 256 + 6 bias chunks x 128 + 8 M9 chunks x 128)')
    for n, v in zip(names[:6], q[:6]):
        print(f'   {n:20s} {v:10d}  {100 * v / q[7]:5.1f}%   {v / q[6]:9.0f} / tile')
m = b[32:36]
print(f'MMA lane: total {m[2]} cycles, wait a_ready {m[0]} ({100 * m[0] / max(1, m[2]):.1f}%), chunks {m[3]}, '
      f'cycles waiting on the weight ring {m[1]} ({100 * m[1] / max(1, m[2]):.1f}%), '
      f'{(m[2] - m[0]) / max(1, m[3]):.0f} cycles per chunk outside a_ready waits (ideal 256)')
for r in range(2):
    w, tot = b[40 + 4 * r], b[41 + 4 * r]
    if tot:
        print(f'weight producer CTA {r}: total {tot} cycles, waiting for free ring stages {w} ({100 * w / tot:.1f}%), '
              f'in expect_tx + cp.async.bulk issue {b[42 + 4 * r]} ({100 * b[42 + 4 * r] / tot:.1f}%)')
if b[49]:
    print(f'weight relay (peer CTA): total {b[49]} cycles, waiting for its own copies {b[48]} ({100 * b[48] / b[49]:.1f}%)')
