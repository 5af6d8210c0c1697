"""Pins oracle/train_batch_oracle.py to the unmodified reference DataPreprocessor.load_cached_next_batch
(golden fixture from oracle/make_golden_train_batch.py): same keys, dtypes, shapes, values bit for bit, over an epoch
boundary (numpy.random.shuffle consumption).  CPU only."""
import numpy
import pytest

from oracle import train_batch_oracle as T
from tests.helpers import load_npz_raw


def unflatten(arrays, prefix):
    out = {}
    for k, v in arrays.items():
        if not k.startswith(prefix):
            continue
        node = out
        parts = k[len(prefix):].split('.')
        for p in parts[:-1]:
            node = node.setdefault(p, {})
        node[parts[-1]] = v.copy()
    return out


def golden_case(name):
    arrays = load_npz_raw('train_batch.npz')
    tables = unflatten(arrays, f'{name}.tables.')
    tables['nerf_data']['resolution'] = tuple(int(x) for x in tables['nerf_data']['resolution'])
    ndc, sparse, num_rays, num_rays_sd, num_gpus = (int(x) for x in arrays[f'{name}.meta'])
    batches = [unflatten(arrays, f'{name}.batch{b}.') for b in range(4)]
    return tables, dict(ndc=bool(ndc), num_rays=num_rays, num_rays_sparse_depth=num_rays_sd if sparse else None,
                        prior_masks=True, prior_weights=False, num_gpus=num_gpus), batches


def assert_batch_equal(got, ref, tag):
    assert set(got) == set(ref), (tag, sorted(got), sorted(ref))
    for k, r in ref.items():
        g = got[k]
        if k == 'common_data':
            assert set(g) == set(r)
            for kk in r:
                assert numpy.array_equal(numpy.asarray(g[kk]), r[kk]), (tag, k, kk)
            continue
        g, r = numpy.asarray(g), numpy.asarray(r)
        assert g.shape == r.shape and g.dtype == r.dtype, (tag, k, g.shape, r.shape, g.dtype, r.dtype)
        assert numpy.array_equal(g, r), (tag, k)


@pytest.mark.parametrize('name', ['llff', 'dtu'])
def test_train_batch_oracle_matches_reference_golden(name):
    tables, kw, batches = golden_case(name)
    state = {'i_batch': 0, 'i_batch_sparse_depth': 0}
    numpy.random.seed(5)
    for b, ref in enumerate(batches):
        got = T.load_cached_next_batch(tables, state, iter_num=100 + b, **kw)
        assert_batch_equal(got, ref, f'{name}/batch{b}')
