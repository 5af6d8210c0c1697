#!/usr/bin/env python
"""Time per training batch: TrainBatchLoaderFused (one upload + one gather launch) vs the reference's statement pattern
on the same GPU (`out = -1 * torch.ones(...).to(device); out[mask] = table[indices[mask]]` per column,
DataPreprocessor01.py:571-724), 4096 + 1024 rays over 3 x 378 x 504 pixels.  Host wall clock with a device sync per
batch (the trainer consumes the batch immediately)."""
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import numpy  # noqa: E402
import torch  # noqa: E402

from oracle.make_golden_train_batch import synthetic_tables  # noqa: E402
from vipnerf_b200.TrainBatchFused01 import TrainBatchLoaderFused  # noqa: E402


def reference_pattern(tables_dev, indices, class_id, device):
    idx = torch.from_numpy(indices).to(device)
    m_nerf = torch.from_numpy(class_id == 1).to(device)
    m_sd = torch.from_numpy(class_id == 2).to(device)
    i_nerf, i_sd = idx[m_nerf], idx[m_sd]
    out = {}
    nd, sd, vp = tables_dev['nerf_data'], tables_dev['sparse_depth_data'], tables_dev['visibility_prior_data']
    for key, tab in (('rays_o', 'rays_o'), ('rays_d', 'rays_d'), ('view_dirs', 'view_dirs'), ('pixel_id', 'pixel_id'),
                     ('target_rgb', 'target_rgb'), ('near', 'near_array'), ('far', 'far_array'), ('rays_o_ndc', 'rays_o_ndc'),
                     ('rays_d_ndc', 'rays_d_ndc'), ('near_ndc', 'near_array_ndc'), ('far_ndc', 'far_array_ndc')):
        t = nd[tab]
        o = (-1 * torch.ones((idx.shape[0],) + tuple(t.shape[1:]), dtype=t.dtype)).to(device)
        o[m_nerf] = t[i_nerf]
        if key != 'target_rgb':
            o[m_sd] = t[i_sd]
        out[key] = o
    for key, tab in (('sparse_depth_values', 'depths'), ('sparse_depth_errors', 'reprojection_errors'), ('sparse_depth_values_ndc', 'depths_ndc')):
        o = (-1 * torch.ones((idx.shape[0], 1))).to(device)
        o[m_sd] = sd[tab][i_sd]
        out[key] = o
    o = (-1 * torch.ones((idx.shape[0], vp['masks'].shape[1]))).to(device)
    o[m_nerf] = vp['masks'][i_nerf]
    out['visibility_prior_masks'] = o
    return out


def main():
    device = torch.device('cuda:0')
    tables = synthetic_tables(3, True, True, n_frames=3, h=378, w=504)
    loader = TrainBatchLoaderFused(tables, device=device, ndc=True, num_rays=4096, num_rays_sparse_depth=1024, prior_masks=True)
    tables_dev = {k: ({kk: (torch.from_numpy(vv).to(device) if isinstance(vv, numpy.ndarray) and kk not in ('indices', 'poses') else vv)
                       for kk, vv in v.items()} if isinstance(v, dict) else v) for k, v in tables.items()}
    res = {}
    for name in ('fused', 'reference_pattern'):
        ts = []
        for it in range(60):
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            if name == 'fused':
                loader.load_cached_next_batch(it, None)
            else:
                idx, cls, _ = loader.select_batch_indices(it, None)
                reference_pattern(tables_dev, idx, cls, device)
            torch.cuda.synchronize()
            ts.append(time.perf_counter() - t0)
        res[name + '_ms_per_batch'] = 1e3 * float(numpy.median(ts[10:]))
    res['speedup'] = res['reference_pattern_ms_per_batch'] / res['fused_ms_per_batch']
    res['rays_per_batch'] = 4096 + 1024
    print(json.dumps(res))


if __name__ == '__main__':
    main()
