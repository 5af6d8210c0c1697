"""GPU parity of the steps either side of the render path: vipnerf_generate_rays and
vipnerf_postprocess_frame through the DataPreprocessorFused plugin, against golden outputs of the UNMODIFIED reference
DataPreprocessor (tests/golden/frame_*.npz) and against the numpy oracle on a full-size frame."""
import numpy
import pytest
import torch

from oracle import frame_oracle as F
from tests.helpers import load_npz_raw
from tests.test_frame_oracle import _model_configs

pytestmark = pytest.mark.gpu

# fp32, same operations as numpy; numpy's stacked 3x3 matmul may fuse multiply-adds, the kernel does not
RAY_TOL = 2e-6


def _plugin(ndc, mc):
    from vipnerf_b200.DataPreprocessorFactory import get_data_preprocessor
    cfg = {'data_loader': {'data_preprocessor_name': 'DataPreprocessorFused01', 'ndc': ndc},
           'model': {'coarse_mlp': {}, 'fine_mlp': {}}, 'device': [0]}
    return get_data_preprocessor(cfg, 'test', model_configs=mc)


def _close(got, ref, key):
    assert tuple(got.shape) == tuple(ref.shape), (key, got.shape, ref.shape)
    scale = max(float(numpy.abs(ref).max()), 1e-30)
    err = float(numpy.abs(got.astype(numpy.float64) - ref.astype(numpy.float64)).max()) / scale
    assert err <= RAY_TOL, (key, err)


@pytest.mark.parametrize('scene', ['fern', 'dtu'])
def test_generate_rays_matches_reference_golden(scene, built_library):
    g = load_npz_raw(f'frame_{scene}.npz')
    mc, ndc = _model_configs(g), bool(g['cfg.ndc'])
    dp = _plugin(ndc, mc)
    a = dp.create_test_data(g['pose.render'])
    b = dp.create_test_data(g['pose.render'], g['pose.view'], list(g['pose.secondary']))
    for tag, got in (('a', a), ('b', b)):
        ref_keys = {k[2:] for k in g if k.startswith(f'{tag}.')}
        assert set(got) == ref_keys, (tag, sorted(got), sorted(ref_keys))
        for k in ref_keys:
            assert got[k].is_cuda and got[k].dtype == torch.float32
            _close(got[k].cpu().numpy(), g[f'{tag}.{k}'], f'{tag}.{k}')
    # constants are exact
    assert torch.equal(a['near'].cpu(), torch.from_numpy(g['a.near']))
    assert torch.equal(b['rays_o2'].cpu(), torch.from_numpy(g['b.rays_o2']))
    assert torch.equal(a['rays_o'].cpu(), torch.from_numpy(g['a.rays_o']))


@pytest.mark.parametrize('scene', ['fern', 'dtu'])
def test_generate_rays_pixel_ranges_tile_the_frame(scene, built_library):
    """A frame generated in ragged pieces (how a frame is sharded over GPUs) equals the frame generated at once;
    an empty range is fine."""
    g = load_npz_raw(f'frame_{scene}.npz')
    mc, ndc = _model_configs(g), bool(g['cfg.ndc'])
    dp = _plugin(ndc, mc)
    whole = dp.create_test_data(g['pose.render'], None, list(g['pose.secondary']))
    R = mc['resolution'][0] * mc['resolution'][1]
    cuts = [0, 1, 130, 130, 517, R]
    parts = [dp.create_test_data(g['pose.render'], None, list(g['pose.secondary']), first_pixel=lo, n_rays=hi - lo)
             for lo, hi in zip(cuts[:-1], cuts[1:])]
    for k in whole:
        assert torch.equal(torch.cat([p[k] for p in parts], dim=0), whole[k]), k
    for part in parts:      # every tensor must be usable as a render input: 16-byte aligned whatever the ray count
        assert all(t.data_ptr() % 16 == 0 for t in part.values())
    with pytest.raises(Exception):
        dp.create_test_data(g['pose.render'], first_pixel=R - 3, n_rays=8)


def test_generate_rays_full_size_frame_against_oracle(built_library):
    """LLFF fern at its real 378 x 504 size: every key against the numpy oracle."""
    g = load_npz_raw('frame_fern.npz')
    mc = _model_configs(g)
    h0, w0 = mc['resolution']
    k = numpy.array(mc['intrinsic'])
    k[0] *= 504 / w0
    k[1] *= 378 / h0
    mc['resolution'], mc['intrinsic'] = [378, 504], k.tolist()
    dp = _plugin(True, mc)
    got = dp.create_test_data(g['pose.render'], g['pose.view'], list(g['pose.secondary']))
    ref = F.create_test_data(mc, True, g['pose.render'], g['pose.view'], list(g['pose.secondary']))
    assert set(got) == set(ref)
    for key in ref:
        _close(got[key].cpu().numpy(), ref[key], key)


@pytest.mark.parametrize('scene', ['fern', 'dtu'])
def test_retrieve_inference_outputs_matches_reference_golden(scene, built_library):
    g = load_npz_raw(f'frame_{scene}.npz')
    mc, ndc = _model_configs(g), bool(g['cfg.ndc'])
    dp = _plugin(ndc, mc)
    outs = {k[4:]: torch.from_numpy(g[k]).cuda() for k in g if k.startswith('net.')}
    ret = dp.retrieve_inference_outputs(outs)
    ref_keys = {k[4:] for k in g if k.startswith('ret.')}
    assert set(ret) == ref_keys
    for k in ref_keys:
        ref = g[f'ret.{k}']
        assert ret[k].dtype == ref.dtype and ret[k].shape == ref.shape, (k, ret[k].dtype, ret[k].shape)
        numpy.testing.assert_array_equal(ret[k], ref, err_msg=k)      # byte / clip / transpose work: bit-exact
