"""CPU restatement (numpy) of the reference's cached training-batch assembly.  TEST INFRASTRUCTURE ONLY.

Follows DataPreprocessor.load_cached_next_batch (src/data_preprocessors/DataPreprocessor01.py:498-530):
select_batch_indices :532-565 (consecutive slices of the pre-shuffled index arrays, numpy.random.shuffle at each epoch
end, class ids 1 = nerf ray / 2 = sparse-depth ray), load_nerf_cached_batch :567-615, load_sparse_depth_cached_batch
:635-683 (the ray columns of the sparse-depth rows are filled too), load_visibility_prior_cached_batch :699-724, and the
num_gpus tiling of common_data :523-529.  Pinned bit for bit to the unmodified reference class by
tests/golden/train_batch.npz (oracle/make_golden_train_batch.py) and, where /root/reference is mounted, live.

`tables` = the reference's preprocessed_data_dict (numpy arrays instead of device tensors); `state` = the loader's
counters ({'i_batch', 'i_batch_sparse_depth'}), updated in place like the reference updates `self`.
"""
from __future__ import annotations

from typing import Dict, Optional

import numpy


def select_batch_indices(tables: dict, state: dict, num_rays: int, num_rays_sparse_depth: Optional[int],
                         image_num: Optional[int] = None):
    """-> (indices int64 [R], class_id uint8 [R]); DataPreprocessor01.py:532-565 (the precrop re-generation at
    iter_num == precrop_iterations, :536-537, is the caller's: it replaces tables['indices'])."""
    sparse = num_rays_sparse_depth is not None and image_num is None
    if image_num is None:
        indices = tables['indices'][state['i_batch']: state['i_batch'] + num_rays]
        state['i_batch'] += num_rays
        if state['i_batch'] >= tables['indices'].size:
            numpy.random.shuffle(tables['indices'])
            state['i_batch'] = 0
    else:
        h, w = tables['nerf_data']['resolution']
        image_index = numpy.where(tables['frame_nums'] == image_num)[0].item()
        indices = numpy.arange(h * w) + (image_index * h * w)
    class_id = numpy.ones_like(indices)
    if sparse:
        sd = tables['sparse_depth_data']['indices']
        indices_sd = sd[state['i_batch_sparse_depth']: state['i_batch_sparse_depth'] + num_rays_sparse_depth]
        state['i_batch_sparse_depth'] += num_rays_sparse_depth
        if state['i_batch_sparse_depth'] >= sd.size:
            numpy.random.shuffle(sd)
            state['i_batch_sparse_depth'] = 0
        indices = numpy.concatenate([indices, indices_sd])
        class_id = numpy.concatenate([class_id, 2 * numpy.ones_like(indices_sd)])
    return indices.astype(numpy.int64), class_id.astype(numpy.uint8), sparse


def column_specs(tables: dict, ndc: bool, sparse: bool, prior_masks: bool, prior_weights: bool):
    """(output key, table, row classes that gather: bit 0 = class 1, bit 1 = class 2) in the reference's key order."""
    nd = tables['nerf_data']
    both = 3 if sparse else 1
    cols = [('rays_o', nd['rays_o'], both), ('rays_d', nd['rays_d'], both), ('view_dirs', nd['view_dirs'], both),
            ('pixel_id', nd['pixel_id'], both), ('target_rgb', nd['target_rgb'], 1), ('near', nd['near_array'], both),
            ('far', nd['far_array'], both)]
    if ndc:
        cols += [('rays_o_ndc', nd['rays_o_ndc'], both), ('rays_d_ndc', nd['rays_d_ndc'], both),
                 ('near_ndc', nd['near_array_ndc'], both), ('far_ndc', nd['far_array_ndc'], both)]
    if sparse:
        sd = tables['sparse_depth_data']
        cols += [('sparse_depth_values', sd['depths'], 2), ('sparse_depth_errors', sd['reprojection_errors'], 2)]
        if ndc:
            cols.append(('sparse_depth_values_ndc', sd['depths_ndc'], 2))
    if prior_masks:
        cols.append(('visibility_prior_masks', tables['visibility_prior_data']['masks'], 1))
    if prior_weights:
        cols.append(('visibility_prior_weights', tables['visibility_prior_data']['weights'], 1))
    return cols


def load_cached_next_batch(tables: dict, state: dict, *, iter_num: int, num_rays: int,
                           num_rays_sparse_depth: Optional[int], ndc: bool, prior_masks: bool, prior_weights: bool,
                           num_gpus: int = 1, image_num: Optional[int] = None) -> Dict[str, object]:
    indices, class_id, sparse = select_batch_indices(tables, state, num_rays, num_rays_sparse_depth, image_num)
    out: Dict[str, object] = {'common_data': {}, 'indices': indices, 'indices_mask_nerf': class_id == 1}
    if sparse:
        out['indices_mask_sparse_depth'] = class_id == 2
    out['iter_num'] = iter_num
    out['num_frames'] = int(tables['frame_nums'].size)
    for key, table, classes in column_specs(tables, ndc, sparse, prior_masks, prior_weights):
        table = numpy.asarray(table)
        take = ((class_id == 1) & bool(classes & 1)) | ((class_id == 2) & bool(classes & 2))
        col = numpy.full((indices.shape[0],) + table.shape[1:], -1, dtype=table.dtype)
        col[take] = table[indices[take]]
        out[key] = col
    if prior_masks or prior_weights:
        poses = numpy.asarray(tables['nerf_data']['poses'])
        out['common_data']['poses'] = numpy.broadcast_to(poses[None], (num_gpus,) + poses.shape).copy()
    return out
