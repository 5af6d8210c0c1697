#!/usr/bin/env python
"""The reference's own PyTorch graph on ONE B200 (cuBLAS / eager kernels, unfused) - BASELINE.md section 3's optional
second baseline - for the eval render and for the training iteration, in fp32 and with torch's allow_tf32, next to
which bench.py's numbers can be read.  Drives the UNMODIFIED reference model staged under oracle/_ref when present
(else the oracle port, which is the same torch graph).  Test / bench infrastructure, not product.

    python tools/torch_gpu_baseline.py > gpurun_out/torch_gpu_baseline.json
"""
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import torch  # noqa: E402

import bench  # noqa: E402
from oracle import build_ref, vipnerf_oracle as O  # noqa: E402
from vipnerf_b200.LossComputerFused01 import LossComputer  # noqa: E402


def ref_model(device):
    cfg = bench.model_configs('bf16')
    cfg['model']['name'] = 'VipNeRF01'
    del cfg['model']['precision']
    if build_ref.ref_available():
        model = build_ref.load_ref_get_model()(cfg, None)
        kind = 'unmodified reference VipNeRF01 (oracle/_ref)'
    else:
        raise SystemExit('oracle/_ref is not staged')
    model.load_state_dict(O.synth_state_dict(0))
    return model.to(device), cfg, kind


def timed(fn, steps, warmup):
    for _ in range(warmup):
        fn()
    torch.cuda.synchronize()
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record()
    for _ in range(steps):
        fn()
    e.record()
    torch.cuda.synchronize()
    return s.elapsed_time(e) / steps


def main():
    device = torch.device('cuda:0')
    R = 4096
    model, cfg, kind = ref_model(device)
    res = {'model': kind, 'rays_per_step': R, 'gpu': torch.cuda.get_device_name(0)}
    eval_batch = {k: v.to(device) for k, v in O.make_rays('fern', R, seed=2).items()}
    train_batch = {k: v.to(device) for k, v in O.make_rays('re10k', R, seed=2, n_sec_views=1).items()}
    sup = {k: (v.to(device) if isinstance(v, torch.Tensor) else v) for k, v in O.make_supervision('re10k', R, 1).items()}
    train_batch.update(sup)
    cfg['losses'] = [{'name': 'MSE01', 'weight': 1}, {'name': 'VisibilityLoss01', 'weight': 0.1},
                     {'name': 'VisibilityPriorLoss01', 'iter_weights': {'0': 0, '30000': 0.001}},
                     {'name': 'SparseDepthMSE01', 'weight': 0.1}]
    computer = LossComputer(cfg)
    opt = torch.optim.Adam(model.parameters(), lr=5e-4)
    for tf32 in (False, True):
        torch.backends.cuda.matmul.allow_tf32 = tf32
        torch.backends.cudnn.allow_tf32 = tf32
        tag = 'tf32' if tf32 else 'fp32'
        model.eval()

        def eval_step():
            with torch.no_grad():
                return model(dict(eval_batch))

        ms = timed(eval_step, 10, 3)
        res[f'eval_{tag}'] = {'ms_per_4096_rays': ms, 'rays_per_s': R / (ms * 1e-3)}
        model.train()

        def train_step():
            opt.zero_grad(set_to_none=True)
            out = model(dict(train_batch))
            computer._compute_torch(train_batch, out, False)['TotalLoss'].backward()
            opt.step()

        torch.cuda.reset_peak_memory_stats()
        ms = timed(train_step, 5, 2)
        res[f'train_{tag}'] = {'ms_per_step': ms, 'rays_per_s': R / (ms * 1e-3),
                               'peak_mem_gb': torch.cuda.max_memory_allocated() / 2 ** 30}
    print(json.dumps(res))


if __name__ == '__main__':
    main()
