"""vipnerf_gather_train_batch / TrainBatchLoaderFused against the golden batches of the unmodified reference
DataPreprocessor.load_cached_next_batch (tests/golden/train_batch.npz): same keys, dtypes, shapes and values bit for
bit, across an epoch boundary, with and without sparse-depth rows; then a larger random case against the numpy oracle."""
import copy
import types

import numpy
import pytest
import torch

from oracle import train_batch_oracle as T
from tests.test_train_batch_oracle import assert_batch_equal, golden_case

pytestmark = pytest.mark.gpu


def _to_numpy(batch):
    out = {}
    for k, v in batch.items():
        if k == 'common_data':
            out[k] = {kk: vv.cpu().numpy() for kk, vv in v.items()}
        elif isinstance(v, torch.Tensor):
            out[k] = v.cpu().numpy()
        else:
            out[k] = numpy.asarray(v)
    return out


@pytest.mark.parametrize('name', ['llff', 'dtu'])
def test_fused_train_batches_match_reference_golden(name, built_library):
    from vipnerf_b200.TrainBatchFused01 import TrainBatchLoaderFused
    tables, kw, batches = golden_case(name)
    loader = TrainBatchLoaderFused(tables, device='cuda:0', **kw)
    numpy.random.seed(5)
    for b, ref in enumerate(batches):
        got = loader.load_cached_next_batch(100 + b, None)
        assert got['rays_o'].is_cuda and got['pixel_id'].dtype == torch.int32 and got['indices_mask_nerf'].dtype == torch.bool
        assert_batch_equal(_to_numpy(got), ref, f'{name}/batch{b}')


def test_fused_train_batches_large_random_and_attach(built_library):
    """4096 + 1024 rays per batch over 3 x 120 x 160 pixels against the numpy oracle, driven through `attach` on an
    object with the reference DataPreprocessor's attributes (whole-image batches via image_num included)."""
    from oracle.make_golden_train_batch import synthetic_tables
    from vipnerf_b200.TrainBatchFused01 import TrainBatchLoaderFused
    tables = synthetic_tables(3, True, True, n_frames=3, h=120, w=160)
    ref_tables = copy.deepcopy(tables)
    dp = types.SimpleNamespace(
        configs={'device': [0], 'data_loader': {'precrop_iterations': -1, 'visibility_prior': {'load_masks': True, 'load_weights': False}}},
        mode='train', ndc=True, use_batching=True, device=torch.device('cuda:0'), i_batch=0, num_rays=4096,
        mip_nerf_used=False, sparse_depth_needed=True, dense_depth_needed=False, visibility_prior_needed=True,
        i_batch_sparse_depth=0, num_rays_sparse_depth=1024, preprocessed_data_dict=tables,
        generate_indices=lambda d, c, it: d['indices'])
    TrainBatchLoaderFused.attach(dp)
    state = {'i_batch': 0, 'i_batch_sparse_depth': 0}
    kw = dict(ndc=True, num_rays=4096, num_rays_sparse_depth=1024, prior_masks=True, prior_weights=False, num_gpus=1)
    for it in range(16):     # 57,600 pixels / 4096: the epoch boundary of both index arrays is crossed
        numpy.random.seed(100 + it)
        got = dp.load_cached_next_batch(it, None)
        numpy.random.seed(100 + it)
        ref = T.load_cached_next_batch(ref_tables, state, iter_num=it, **kw)
        assert_batch_equal(_to_numpy(got), ref, f'iter{it}')
    got = dp.load_cached_next_batch(99, 7)       # a whole training image (frame number 7 = image index 1)
    ref = T.load_cached_next_batch(ref_tables, state, iter_num=99, image_num=7, **kw)
    assert got['rays_o'].shape == (120 * 160, 3) and 'indices_mask_sparse_depth' not in got
    assert_batch_equal(_to_numpy(got), ref, 'image')
