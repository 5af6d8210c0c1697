"""Training batches assembled on the GPU in one launch: the caller side of the training step.

Reference being replaced: `DataPreprocessor.load_cached_next_batch` (src/data_preprocessors/DataPreprocessor01.py
:498-530 with :567-615, :635-683, :699-724) - per batch ~40 statements of the form
`out = -1 * torch.ones(...); out[mask] = table[indices[mask]]`, each a boolean-index (device sync) plus several small
launches.  Here: the reference's index selection on the host (`select_batch_indices` :532-565, same numpy RNG calls, so
a seeded run visits the same pixels), ONE upload of indices + row classes, ONE `vipnerf_gather_train_batch` launch that
fills every column.  Same keys, dtypes, shapes and values (bit for bit) as the reference's batch dict.

    loader = TrainBatchLoaderFused.attach(train_data_preprocessor)   # a constructed reference DataPreprocessor (mode 'train')
    batch = train_data_preprocessor.get_next_batch(iter_num)         # now served by the fused loader

`attach` reads the reference object's per-pixel caches (`preprocessed_data_dict`, already on the device) and replaces its
`load_cached_next_batch`; building those caches (image loading, COLMAP depths, prior files; :60-487) stays the
reference's.  mip-NeRF radii and dense depth are not part of the shipped ViP-NeRF configs and raise NotImplementedError.
"""
from __future__ import annotations

import ctypes
from typing import Dict, Optional

import numpy
import torch

from . import _lib


def _dev(t, device) -> torch.Tensor:
    if isinstance(t, numpy.ndarray):
        t = torch.from_numpy(t)
    t = t.to(device)
    if t.dtype not in (torch.float32, torch.int32):
        raise TypeError(f'per-pixel caches must be float32 / int32, got {t.dtype}')
    return t.contiguous()


class TrainBatchLoaderFused:
    def __init__(self, tables: dict, *, device, ndc: bool, num_rays: int, num_rays_sparse_depth: Optional[int] = None,
                 prior_masks: bool = False, prior_weights: bool = False, num_gpus: int = 1,
                 precrop_iterations: int = -1, regenerate_indices=None):
        """`tables`: the reference's preprocessed_data_dict (nerf_data / sparse_depth_data / visibility_prior_data
        per-pixel caches as torch tensors or numpy arrays; `indices` arrays as numpy, shuffled in place like the
        reference does).  `regenerate_indices(iter_num)`: the reference's generate_indices at the end of pre-cropping."""
        if not torch.cuda.is_available():
            raise RuntimeError('TrainBatchLoaderFused needs a CUDA device; there is no CPU fallback')
        self.device = torch.device(device)
        self.tables, self.ndc = tables, ndc
        self.num_rays, self.num_rays_sparse_depth = num_rays, num_rays_sparse_depth
        self.prior_masks, self.prior_weights, self.num_gpus = prior_masks, prior_weights, num_gpus
        self.precrop_iterations, self.regenerate_indices = precrop_iterations, regenerate_indices
        self.i_batch, self.i_batch_sparse_depth = 0, 0
        nd = tables['nerf_data']
        both_sparse = num_rays_sparse_depth is not None
        dev = lambda t: _dev(t, self.device)
        # (output key, device table, row classes that gather when the batch has sparse-depth rows / when it has none)
        self.columns = [('rays_o', dev(nd['rays_o']), 3), ('rays_d', dev(nd['rays_d']), 3), ('view_dirs', dev(nd['view_dirs']), 3),
                        ('pixel_id', dev(nd['pixel_id']), 3), ('target_rgb', dev(nd['target_rgb']), 1),
                        ('near', dev(nd['near_array']), 3), ('far', dev(nd['far_array']), 3)]
        if ndc:
            self.columns += [('rays_o_ndc', dev(nd['rays_o_ndc']), 3), ('rays_d_ndc', dev(nd['rays_d_ndc']), 3),
                             ('near_ndc', dev(nd['near_array_ndc']), 3), ('far_ndc', dev(nd['far_array_ndc']), 3)]
        self.sparse_columns = []
        if both_sparse:
            sd = tables['sparse_depth_data']
            self.sparse_columns = [('sparse_depth_values', dev(sd['depths']), 2), ('sparse_depth_errors', dev(sd['reprojection_errors']), 2)]
            if ndc:
                self.sparse_columns.append(('sparse_depth_values_ndc', dev(sd['depths_ndc']), 2))
        self.prior_columns = []
        if prior_masks:
            self.prior_columns.append(('visibility_prior_masks', dev(tables['visibility_prior_data']['masks']), 1))
        if prior_weights:
            self.prior_columns.append(('visibility_prior_weights', dev(tables['visibility_prior_data']['weights']), 1))
        self.poses = None
        if prior_masks or prior_weights:
            poses = nd['poses']
            self.poses = (torch.from_numpy(poses) if isinstance(poses, numpy.ndarray) else poses).to(self.device)

    # ------------------------------------------------------------------ DataPreprocessor01.select_batch_indices :532-565
    def select_batch_indices(self, iter_num: int, image_num: Optional[int]):
        t = self.tables
        sparse = self.num_rays_sparse_depth is not None and image_num is None
        if image_num is None:
            if iter_num == self.precrop_iterations and self.regenerate_indices is not None:
                t['indices'] = self.regenerate_indices(iter_num)
            indices = t['indices'][self.i_batch: self.i_batch + self.num_rays]
            self.i_batch += self.num_rays
            if self.i_batch >= t['indices'].size:
                numpy.random.shuffle(t['indices'])
                self.i_batch = 0
        else:
            h, w = t['nerf_data']['resolution']
            image_index = numpy.where(numpy.asarray(t['frame_nums']) == image_num)[0].item()
            indices = numpy.arange(h * w) + (image_index * h * w)
        class_id = numpy.ones(indices.shape[0], dtype=numpy.uint8)
        if sparse:
            sd = t['sparse_depth_data']['indices']
            indices_sd = sd[self.i_batch_sparse_depth: self.i_batch_sparse_depth + self.num_rays_sparse_depth]
            self.i_batch_sparse_depth += self.num_rays_sparse_depth
            if self.i_batch_sparse_depth >= sd.size:
                numpy.random.shuffle(sd)
                self.i_batch_sparse_depth = 0
            indices = numpy.concatenate([indices, indices_sd])
            class_id = numpy.concatenate([class_id, numpy.full(indices_sd.shape[0], 2, dtype=numpy.uint8)])
        return indices.astype(numpy.int64), class_id, sparse

    # ------------------------------------------------------------------ DataPreprocessor01.load_cached_next_batch :498-530
    def load_cached_next_batch(self, iter_num: int, image_num: Optional[int] = None) -> Dict[str, object]:
        lib = _lib.load()
        indices, class_id, sparse = self.select_batch_indices(iter_num, image_num)
        R = indices.shape[0]
        # one pinned staging buffer, one upload: [indices int64 | class ids uint8]
        host = torch.empty(R * 9, dtype=torch.uint8).pin_memory()
        host[:R * 8].view(torch.int64).copy_(torch.from_numpy(indices))
        host[R * 8:].copy_(torch.from_numpy(class_id))
        dev = host.to(self.device, non_blocking=True)
        indices_dev, class_dev = dev[:R * 8].view(torch.int64), dev[R * 8:]
        batch: Dict[str, object] = {'common_data': {}, 'indices': indices_dev, 'indices_mask_nerf': class_dev == 1}
        if sparse:
            batch['indices_mask_sparse_depth'] = class_dev == 2
        batch['iter_num'] = iter_num
        batch['num_frames'] = int(numpy.asarray(self.tables['frame_nums']).size)
        specs = self.columns + (self.sparse_columns if sparse else []) + self.prior_columns
        cols = (_lib.GatherColumn * len(specs))()
        for i, (key, table, classes) in enumerate(specs):
            if not sparse:
                classes &= 1
            out = torch.empty((R,) + tuple(table.shape[1:]), dtype=table.dtype, device=self.device)
            width = 1
            for d in table.shape[1:]:
                width *= int(d)
            cols[i] = _lib.GatherColumn(table.data_ptr(), out.data_ptr(), width, classes, int(table.dtype == torch.int32), 0)
            batch[key] = out
        with torch.cuda.device(self.device):
            _lib.check(lib.vipnerf_gather_train_batch(indices_dev.data_ptr(), class_dev.data_ptr(), R, cols, len(specs),
                                                      torch.cuda.current_stream(self.device).cuda_stream),
                       'vipnerf_gather_train_batch')
        if self.poses is not None:   # common data is tiled per GPU for nn.DataParallel's scatter (:523-529)
            batch['common_data']['poses'] = self.poses[None].repeat([self.num_gpus] + [1] * self.poses.ndim)
        return batch

    # ------------------------------------------------------------------ binding to a reference DataPreprocessor object
    @classmethod
    def attach(cls, dp) -> 'TrainBatchLoaderFused':
        """Replaces `dp.load_cached_next_batch` of a constructed reference DataPreprocessor (mode 'train', batching on)."""
        if getattr(dp, 'mip_nerf_used', False) or getattr(dp, 'dense_depth_needed', False):
            raise NotImplementedError('mip-NeRF radii / dense depth batches are not part of this build')
        if not dp.use_batching:
            raise NotImplementedError('only the cached (batching=True) loader is replaced')
        cfg = dp.configs['data_loader']
        prior = cfg.get('visibility_prior', {}) if dp.visibility_prior_needed and dp.mode == 'train' else {}
        sparse = dp.sparse_depth_needed and dp.mode == 'train'
        loader = cls(dp.preprocessed_data_dict, device=dp.device, ndc=dp.ndc, num_rays=dp.num_rays,
                     num_rays_sparse_depth=dp.num_rays_sparse_depth if sparse else None,
                     prior_masks=bool(prior.get('load_masks', False)), prior_weights=bool(prior.get('load_weights', False)),
                     num_gpus=len(dp.configs['device']), precrop_iterations=cfg.get('precrop_iterations', -1),
                     regenerate_indices=lambda it: dp.generate_indices(dp.preprocessed_data_dict, None, it))
        loader.i_batch = dp.i_batch
        loader.i_batch_sparse_depth = getattr(dp, 'i_batch_sparse_depth', 0)
        dp.load_cached_next_batch = loader.load_cached_next_batch
        return loader
