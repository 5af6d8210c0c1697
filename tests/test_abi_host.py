"""Host-side checks that need no GPU: the C-ABI library builds, loads and exports every symbol the header
declares; configuration errors and the plugin's host logic behave like the reference's."""
import ctypes
import os
import re

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    text = open(os.path.join(ROOT, 'include', 'vipnerf.h')).read()
    text = re.sub(r'/\*.*?\*/', '', text, flags=re.S)
    return sorted(set(re.findall(r'\b(vipnerf_[a-z0-9_]+)\s*\(', text)))


def test_header_symbols_exported(built_library):
    lib = ctypes.CDLL(built_library)
    names = declared_symbols()
    assert 'vipnerf_render_forward' in names and len(names) >= 10
    for name in names:
        assert hasattr(lib, name), f'{name} declared in include/vipnerf.h but not exported'


def test_binding_covers_header(built_library):
    from vipnerf_b200 import _lib
    assert sorted(_lib.EXPORTS) == declared_symbols()
    lib = _lib.load()
    assert lib.vipnerf_abi_version() == _lib.ABI_VERSION


def test_struct_sizes_match_header(built_library):
    from vipnerf_b200 import _lib
    assert ctypes.sizeof(_lib.Cfg) == 48
    assert ctypes.sizeof(_lib.Rays) == 14 * 8
    assert ctypes.sizeof(_lib.PassOut) == 15 * 8
    assert ctypes.sizeof(_lib.Out) == 30 * 8


def test_config_validation(built_library):
    from vipnerf_b200 import _lib
    lib = _lib.load()
    ok = _lib.make_cfg(precision='bf16')
    assert lib.vipnerf_check_config(ctypes.byref(ok)) == 0
    assert lib.vipnerf_packed_weight_bytes(ctypes.byref(ok)) == 27648 + (64 + 6) * 16384 + 8192   # 68 weight chunks (8 half-size; feature_linear is folded into M9) + 6 bias chunks + the view-direction chunk
    assert lib.vipnerf_packed_weight_bytes(ctypes.byref(_lib.make_cfg(precision='fp16'))) == 27648 + (64 + 6) * 16384 + 8192
    x3 = _lib.make_cfg(precision='bf16x3')
    assert lib.vipnerf_packed_weight_bytes(ctypes.byref(x3)) == 27648 + 2 * ((64 + 6) * 16384 + 8192)
    f32 = _lib.make_cfg(precision='fp32')
    assert lib.vipnerf_packed_weight_bytes(ctypes.byref(f32)) == 27648 + (589824 + 589824) * (4 + 2) + 128 * 64 * 2   # forward images + [out][in] images (backward-data chain, tensor-core forward), + their fp16 mirror and the fp16 view-direction columns (fp16 training mode)
    # training buffers: fp32 only; about 11 KB of saved activations per sample point
    assert lib.vipnerf_train_saved_bytes(ctypes.byref(ok), 4096) == 0
    f32v = _lib.make_cfg(precision='fp32', n_sec_views=1)
    per_point = lib.vipnerf_train_saved_bytes(ctypes.byref(f32v), 4096) / (4096 * 256)
    assert 10500 < per_point < 11500
    f16v = _lib.make_cfg(precision='fp32', n_sec_views=1, train_precision='fp16')   # fp16 arrays: half of everything
    assert 5250 < lib.vipnerf_train_saved_bytes(ctypes.byref(f16v), 4096) / (4096 * 256) < 5800
    both = _lib.make_cfg(precision='fp32', train_precision='fp16')
    both.flags |= _lib.FLAG_TRAIN_TF32                                              # the two tensor-core modes are exclusive
    assert lib.vipnerf_train_saved_bytes(ctypes.byref(both), 4096) == 0
    assert lib.vipnerf_train_workspace_bytes(ctypes.byref(f32v), 4096) > 4096 * 192 * 9 * 1024
    assert lib.vipnerf_train_forward(ctypes.byref(ok), None, 16, None, None, None, None, None, None, 0, None, 0, None) == -2
    assert lib.vipnerf_train_forward(ctypes.byref(f32), None, 16, None, None, None, None, None, None, 0, None, 0, None) == -1
    assert lib.vipnerf_train_backward(ctypes.byref(f32), None, 16, None, None, None, None, None, 0, None, None, None, 0, None) == -1
    assert lib.vipnerf_workspace_bytes(ctypes.byref(ok), 4096) > 4096 * (64 + 192 * 6) * 4
    bad = _lib.make_cfg(width=128)
    assert lib.vipnerf_check_config(ctypes.byref(bad)) == -2
    assert b'W=128' in lib.vipnerf_last_error()
    with pytest.raises(NotImplementedError):
        _lib.check(lib.vipnerf_check_config(ctypes.byref(bad)), 'check')
    bad_abi = _lib.make_cfg()
    bad_abi.abi = 99
    assert lib.vipnerf_check_config(ctypes.byref(bad_abi)) == -5
    sec_tc = _lib.make_cfg(precision='bf16', n_sec_views=2)     # secondary views run on the tensor path (<= 8 per tile)
    assert lib.vipnerf_check_config(ctypes.byref(sec_tc)) == 0
    sec_tc = _lib.make_cfg(precision='bf16', n_sec_views=9)
    assert lib.vipnerf_check_config(ctypes.byref(sec_tc)) == -2
    assert lib.vipnerf_check_config(None) == -1
    # NULL arguments are reported, never dereferenced
    assert lib.vipnerf_render_forward(ctypes.byref(ok), None, 16, None, None, None, None, 0, None) == -1


def _configs(ndc=True, **model_over):
    mlp = dict(num_samples=64, netdepth=8, netwidth=256, points_positional_encoding_degree=10,
               views_positional_encoding_degree=4, use_view_dirs=True, view_dependent_rgb=True, predict_visibility=True)
    model = dict(name='VipNeRFFused01', coarse_mlp=dict(mlp), fine_mlp=dict(mlp, num_samples=128), chunk=4096,
                 lindisp=False, netchunk=16384, perturb=True, raw_noise_std=1.0, white_bkgd=False)
    model.update(model_over)
    return {'data_loader': {'ndc': ndc}, 'model': model}


def test_plugin_factory_and_state_dict():
    from oracle import vipnerf_oracle as O
    from vipnerf_b200.ModelFactory import get_model
    model = get_model(_configs(), None)
    assert type(model).__name__ == 'VipNeRFFused'
    sd = O.synth_state_dict(0)
    assert set(model.state_dict().keys()) == set(sd.keys())          # reference checkpoint keys
    model.load_state_dict(sd)
    assert sum(p.numel() for p in model.parameters()) == 1191946
    # DataParallel checkpoints carry a 'module.' prefix (Trainer01.py:357-362)
    wrapped = torch.nn.DataParallel(model)
    wrapped.load_state_dict({f'module.{k}': v for k, v in sd.items()})
    with pytest.raises(RuntimeError):
        get_model({'data_loader': {'ndc': True}, 'model': {'name': 'NoSuchModel01'}}, None)


def test_plugin_rejects_cpu_tensors_and_unsupported_shapes():
    from oracle import vipnerf_oracle as O
    from vipnerf_b200.ModelFactory import get_model
    model = get_model(_configs(ndc=False), None).eval()
    batch = O.make_rays('dtu', 8)
    with pytest.raises(RuntimeError, match='no CPU fallback'):
        model(batch)
    model.train()   # training mode has no CPU path either
    with pytest.raises(RuntimeError, match='no CPU fallback'):
        model(O.make_rays('dtu', 8, n_sec_views=2))
    with pytest.raises(NotImplementedError):
        get_model(_configs(coarse_mlp=dict(_configs()['model']['coarse_mlp'], netwidth=128)), None)


def test_missing_library_fails_loudly(monkeypatch, tmp_path):
    from vipnerf_b200 import _lib
    monkeypatch.setattr(_lib, '_lib', None)
    monkeypatch.setattr(_lib, 'LIB_PATH', str(tmp_path / 'libvipnerf_b200.so'))
    with pytest.raises(_lib.VipNeRFLibraryError, match='no CPU fallback'):
        _lib.load()


def test_training_modes_host_logic():
    """The training arithmetic is a config key and a cfg flag: 'fp32' / 'tf32' / 'fp16' map to the header's flags, anything
    else is refused where the reference would refuse an unknown config value; GraphedTrainStep refuses the set-ups it
    cannot capture before it touches the device; the bench's algorithmic byte counts are the sums DESIGN.md derives."""
    import bench
    from vipnerf_b200 import _lib, training
    from vipnerf_b200.ModelFactory import get_model
    header = open(os.path.join(ROOT, 'include', 'vipnerf.h')).read()
    for name, value in (('VIPNERF_FLAG_TRAIN_TF32', _lib.FLAG_TRAIN_TF32), ('VIPNERF_FLAG_TRAIN_F16', _lib.FLAG_TRAIN_F16)):
        m = re.search(rf'#define\s+{name}\s+\(1u << (\d+)\)', header)
        assert m and (1 << int(m.group(1))) == value, name
    assert _lib.make_cfg(precision='fp32', train_precision='fp16').flags & _lib.FLAG_TRAIN_F16
    assert _lib.make_cfg(precision='fp32', train_tf32=True).flags & _lib.FLAG_TRAIN_TF32        # the older spelling
    assert _lib.make_cfg(precision='fp32').flags & (_lib.FLAG_TRAIN_TF32 | _lib.FLAG_TRAIN_F16) == 0
    with pytest.raises(ValueError):
        _lib.make_cfg(train_precision='bf16')
    for tp in ('fp32', 'tf32', 'fp16'):
        get_model(_configs(train_precision=tp), None)
    with pytest.raises(ValueError):
        get_model(_configs(train_precision='fp8'), None)
    model = get_model(_configs(train_precision='fp16'), None)          # rng defaults to the reference's CPU draws
    opt = torch.optim.Adam(model.parameters(), capturable=True)
    with pytest.raises(ValueError, match='rng'):
        training.GraphedTrainStep(model, None, opt, {})
    model = get_model(_configs(train_precision='fp16', rng='device'), None)
    with pytest.raises(ValueError, match='capturable'):
        training.GraphedTrainStep(model, None, torch.optim.Adam(model.parameters()), {})
    assert bench.TRAIN_F16_BYTES_PER_POINT == 34_160 and bench.TRAIN_TC_BYTES_PER_POINT == 82_652


def test_gradient_scale_rule(built_library):
    """vipnerf_grad_scale = the device function that scales the fp16 gradient arrays: a power of two that moves the
    measured maximum into [16, 32) (11 binades below fp16's 65504), 1 for the values no maximum can define a scale from."""
    import math
    from vipnerf_b200 import _lib
    lib = _lib.load()
    assert lib.vipnerf_grad_scale(20.0) == 1.0 and lib.vipnerf_grad_scale(16.0) == 1.0 and lib.vipnerf_grad_scale(31.9) == 1.0
    assert lib.vipnerf_grad_scale(1.0) == 16.0 and lib.vipnerf_grad_scale(32.0) == 0.5
    for bad in (0.0, -3.0, float('inf'), float('nan')):
        assert lib.vipnerf_grad_scale(bad) == 1.0
    g = torch.Generator().manual_seed(0)
    for amax in torch.exp(torch.empty(200).uniform_(-60.0, 60.0, generator=g)).tolist():
        s = lib.vipnerf_grad_scale(amax)
        assert s > 0 and math.log2(s) == round(math.log2(s)), (amax, s)       # a power of two: exact to apply and to undo
        assert 16.0 <= float(torch.tensor(amax, dtype=torch.float32)) * s < 32.0, (amax, s)
    assert lib.vipnerf_grad_scale(1e-45) == 2.0 ** 120                        # subnormal maxima: the exponent is clamped
