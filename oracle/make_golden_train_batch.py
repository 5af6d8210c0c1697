"""Generates tests/golden/train_batch.npz from the UNMODIFIED reference DataPreprocessor (build container only):
load_cached_next_batch (src/data_preprocessors/DataPreprocessor01.py:498-530) on seeded synthetic per-pixel caches, three
consecutive batches across an epoch boundary (so the numpy.random.shuffle consumption is in the fixture), LLFF-style
(NDC + sparse depth + visibility-prior masks) and DTU-style (world space, no sparse depth).

    python -m oracle.make_golden_train_batch
"""
import os

import numpy
import torch

from oracle import ref_loader

HERE = os.path.dirname(os.path.abspath(__file__))
OUT = os.path.join(os.path.dirname(HERE), 'tests', 'golden', 'train_batch.npz')


def synthetic_tables(seed: int, ndc: bool, sparse: bool, n_frames: int = 3, h: int = 6, w: int = 8):
    rng = numpy.random.default_rng(seed)
    n = n_frames * h * w
    f32 = lambda *s: rng.standard_normal(s).astype(numpy.float32)
    nd = {'resolution': (h, w), 'rays_o': f32(n, 3), 'rays_d': f32(n, 3), 'view_dirs': f32(n, 3),
          'pixel_id': rng.integers(0, 100, size=(n, 3)).astype(numpy.int32), 'target_rgb': rng.random((n, 3)).astype(numpy.float32),
          'near_array': numpy.full((n, 1), 1.0, numpy.float32), 'far_array': numpy.full((n, 1), 6.2, numpy.float32),
          'poses': f32(n_frames, 3, 4)}
    if ndc:
        nd.update({'rays_o_ndc': f32(n, 3), 'rays_d_ndc': f32(n, 3), 'near_array_ndc': numpy.zeros((n, 1), numpy.float32),
                   'far_array_ndc': numpy.ones((n, 1), numpy.float32)})
    tables = {'nerf_data': nd, 'frame_nums': numpy.arange(n_frames) * 7, 'indices': rng.permutation(n),
              'visibility_prior_data': {'masks': (rng.random((n, n_frames - 1)) > 0.4).astype(numpy.float32)}}
    if sparse:
        sd_idx = rng.permutation(n)[: n // 5]
        depths = numpy.full((n, 1), -1, numpy.float32)
        depths[sd_idx] = rng.random((sd_idx.size, 1)).astype(numpy.float32) * 5 + 1
        tables['sparse_depth_data'] = {'indices': sd_idx.copy(), 'depths': depths, 'reprojection_errors': f32(n, 1),
                                       'depths_ndc': f32(n, 1)}
    return tables


def flatten(prefix, d, out):
    for k, v in d.items():
        if isinstance(v, dict):
            flatten(f'{prefix}{k}.', v, out)
        else:
            out[f'{prefix}{k}'] = numpy.asarray(v)


def reference_loader(tables, ndc, sparse, num_rays, num_rays_sd, num_gpus):
    """The reference class without its dataset-reading constructor: attributes set as __init__ :18-41 would."""
    ref_loader.import_reference()
    from data_preprocessors.DataPreprocessor01 import DataPreprocessor
    dp = object.__new__(DataPreprocessor)
    dp.configs = {'device': list(range(num_gpus)), 'data_loader': {'precrop_iterations': -1, 'visibility_prior': {'load_masks': True, 'load_weights': False}}}
    dp.mode, dp.ndc, dp.use_batching, dp.device = 'train', ndc, True, torch.device('cpu')
    dp.i_batch, dp.num_rays = 0, num_rays
    dp.mip_nerf_used, dp.sparse_depth_needed, dp.dense_depth_needed, dp.visibility_prior_needed = False, sparse, False, True
    if sparse:
        dp.i_batch_sparse_depth, dp.num_rays_sparse_depth = 0, num_rays_sd

    def to_torch(d):
        return {k: (to_torch(v) if isinstance(v, dict) else (torch.from_numpy(v) if isinstance(v, numpy.ndarray) and k not in ('indices', 'frame_nums', 'poses') else v))
                for k, v in d.items()}
    dp.preprocessed_data_dict = to_torch(tables)
    return dp


def main():
    arrays = {}
    for name, ndc, sparse, num_rays, num_rays_sd, num_gpus in (('llff', True, True, 50, 12, 2), ('dtu', False, False, 64, None, 1)):
        tables = synthetic_tables(11 if ndc else 12, ndc, sparse)
        flatten(f'{name}.tables.', tables, arrays)
        dp = reference_loader({k: (dict(v) if isinstance(v, dict) else (v.copy() if isinstance(v, numpy.ndarray) else v)) for k, v in tables.items()},
                              ndc, sparse, num_rays, num_rays_sd, num_gpus)
        dp.preprocessed_data_dict['indices'] = tables['indices'].copy()
        if sparse:
            dp.preprocessed_data_dict['sparse_depth_data']['indices'] = tables['sparse_depth_data']['indices'].copy()
        numpy.random.seed(5)
        for b in range(4):     # 4 x 50 rays over 144 pixels: the third batch ends the epoch (shuffle), the fourth follows it
            batch = dp.load_cached_next_batch(100 + b, None)
            for k, v in batch.items():
                if k == 'common_data':
                    arrays[f'{name}.batch{b}.common_data.poses'] = v['poses'].numpy().copy()
                elif isinstance(v, torch.Tensor):
                    arrays[f'{name}.batch{b}.{k}'] = v.numpy().copy()   # on the CPU `indices` aliases the array the next shuffle permutes
                else:
                    arrays[f'{name}.batch{b}.{k}'] = numpy.asarray(v)
        arrays[f'{name}.meta'] = numpy.array([int(ndc), int(sparse), num_rays, num_rays_sd or 0, num_gpus])
    numpy.savez_compressed(OUT, **arrays)
    print(OUT, len(arrays), 'arrays', os.path.getsize(OUT), 'bytes')


if __name__ == '__main__':
    main()
