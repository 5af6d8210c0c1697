"""The frame oracle (oracle/frame_oracle.py: ray generation + output post-processing) against golden outputs of the
UNMODIFIED reference DataPreprocessor (tests/golden/frame_*.npz, made by oracle/make_golden_frames.py), and the host
logic of the DataPreprocessorFused plugin that needs no GPU."""
import numpy
import pytest

from oracle import frame_oracle as F
from tests.helpers import load_npz_raw


def _model_configs(g):
    return {'resolution': [int(v) for v in g['cfg.resolution']], 'intrinsic': g['cfg.intrinsic'].tolist(),
            'average_pose': g['cfg.average_pose'].tolist(), 'translation_scale': g['cfg.translation_scale'].item(),
            'near': g['cfg.near'].item(), 'far': g['cfg.far'].item(), 'near_ndc': g['cfg.near_ndc'].item(),
            'far_ndc': g['cfg.far_ndc'].item()}


@pytest.mark.parametrize('scene', ['fern', 'dtu'])
def test_create_test_data_matches_reference(scene):
    g = load_npz_raw(f'frame_{scene}.npz')
    mc, ndc = _model_configs(g), bool(g['cfg.ndc'])
    a = F.create_test_data(mc, ndc, g['pose.render'])
    b = F.create_test_data(mc, ndc, g['pose.render'], g['pose.view'], list(g['pose.secondary']))
    for tag, got in (('a', a), ('b', b)):
        ref_keys = {k[2:] for k in g if k.startswith(f'{tag}.')}
        assert set(got) == ref_keys, (tag, sorted(got), sorted(ref_keys))
        for k in ref_keys:
            ref = g[f'{tag}.{k}']
            assert got[k].shape == ref.shape and got[k].dtype == ref.dtype, (tag, k, got[k].shape, ref.shape)
            numpy.testing.assert_array_equal(got[k], ref, err_msg=f'{tag}.{k}')   # same numpy ops: bit-identical


@pytest.mark.parametrize('scene', ['fern', 'dtu'])
def test_retrieve_inference_outputs_matches_reference(scene):
    g = load_npz_raw(f'frame_{scene}.npz')
    ndc = bool(g['cfg.ndc'])
    outs = {k[4:]: g[k] for k in g if k.startswith('net.')}
    ret = F.retrieve_inference_outputs(outs, [int(v) for v in g['cfg.resolution']], ndc)
    ref_keys = {k[4:] for k in g if k.startswith('ret.')}
    assert set(ret) == ref_keys
    for k in ref_keys:
        assert ret[k].dtype == g[f'ret.{k}'].dtype, k
        numpy.testing.assert_array_equal(ret[k], g[f'ret.{k}'], err_msg=k)
    # the fixture contains exact .5 cases: numpy.round is half-to-even
    x = (numpy.arange(256, dtype=numpy.float32) + 0.5) / 255.0
    assert ret['image'].reshape(-1, 3)[:256, 0].tolist() == numpy.round(numpy.clip(x, 0, 1) * 255).astype('uint8').tolist()


def test_plugin_pose_preprocessing_matches_oracle():
    """The plugin's host-side pose algebra is its own restatement; it must equal the oracle's bit for bit."""
    from vipnerf_b200.DataPreprocessorFused01 import preprocess_test_poses
    g = load_npz_raw('frame_fern.npz')
    poses = numpy.stack([g['pose.render'], g['pose.view'], *g['pose.secondary']])
    sc, avg = g['cfg.translation_scale'].item(), g['cfg.average_pose']
    numpy.testing.assert_array_equal(preprocess_test_poses(poses, sc, avg), F.preprocess_test_poses(poses, sc, avg))


def test_plugin_rejects_what_it_does_not_cover():
    from vipnerf_b200.DataPreprocessorFactory import get_data_preprocessor
    cfg = {'data_loader': {'data_preprocessor_name': 'DataPreprocessorFused01', 'ndc': True}, 'model': {}, 'device': None}
    with pytest.raises(NotImplementedError):
        get_data_preprocessor(cfg, 'train', model_configs={})
    with pytest.raises(RuntimeError):
        get_data_preprocessor(dict(cfg, data_loader=dict(cfg['data_loader'], data_preprocessor_name='Nope01')), 'test')
