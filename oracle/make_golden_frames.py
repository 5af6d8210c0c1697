"""Generates tests/golden/frame_*.npz by running the UNMODIFIED reference DataPreprocessor
(/root/reference/src/data_preprocessors/DataPreprocessor01.py) in test mode: create_test_data :776-864 for real
camera poses of the committed trajectories, and retrieve_inference_outputs :866-894 on synthetic network outputs.
Run in the build container only:  python oracle/make_golden_frames.py
TEST INFRASTRUCTURE ONLY.  Frames are rendered at a small resolution (intrinsics scaled accordingly) to keep the
fixtures small; the arithmetic is resolution independent.
"""
from __future__ import annotations

import json
import os
import sys

import numpy
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import ref_loader  # noqa: E402

GOLDEN = os.path.join(ROOT, 'tests', 'golden')
REF = ref_loader.REFERENCE_ROOT

CASES = {
    # scene: (training run, scene id, pose csv, (h, w), pose rows (render, view, secondary...))
    'fern': ('train0012', 'fern', 'data/databases/NeRF_LLFF/data/train_test_sets/set03/video_poses01/fern.csv', (27, 36)),
    'dtu': ('train0042', '00008', 'data/databases/DTU/data/train_test_sets/set03/video_poses01/00008.csv', (30, 40)),   # csv absent: synthetic poses
}


def small_model_configs(run, scene, hw):
    with open(os.path.join(REF, 'runs', 'training', run, scene, 'ModelConfigs.json')) as f:
        mc = json.load(f)
    h0, w0 = mc['resolution']
    h, w = hw
    k = numpy.array(mc['intrinsic'], dtype=numpy.float64)
    k[0] *= w / w0
    k[1] *= h / h0
    mc['resolution'] = [h, w]
    mc['intrinsic'] = k.tolist()
    return mc


def synthetic_poses(n):
    """World-to-camera matrices on a small arc around the scene (deterministic)."""
    out = []
    for i in range(n):
        a, b = 0.35 * numpy.sin(0.13 * i), 0.25 * numpy.cos(0.07 * i)
        ry = numpy.array([[numpy.cos(a), 0, numpy.sin(a)], [0, 1, 0], [-numpy.sin(a), 0, numpy.cos(a)]])
        rx = numpy.array([[1, 0, 0], [0, numpy.cos(b), -numpy.sin(b)], [0, numpy.sin(b), numpy.cos(b)]])
        m = numpy.eye(4)
        m[:3, :3] = ry @ rx
        m[:3, 3] = [0.4 * numpy.sin(0.11 * i), 0.3 * numpy.cos(0.05 * i), 1.2 + 0.2 * numpy.sin(0.09 * i)]
        out.append(m)
    return numpy.stack(out)


def main():
    ref_loader.import_reference()
    from data_preprocessors.DataPreprocessorFactory import get_data_preprocessor
    g = numpy.random.default_rng(7)
    for name, (run, scene, csv, hw) in CASES.items():
        with open(os.path.join(REF, 'runs', 'training', run, 'Configs.json')) as f:
            cfg = json.load(f)
        cfg['device'] = None
        ndc = cfg['data_loader']['ndc']
        mc = small_model_configs(run, scene, hw)
        csv_path = os.path.join(REF, csv)
        if os.path.isfile(csv_path):
            poses = numpy.loadtxt(csv_path, delimiter=',').reshape(-1, 4, 4)
        else:
            poses = synthetic_poses(100)   # the DTU trajectories are not committed in the reference repo
        dp = get_data_preprocessor(cfg, mode='test', model_configs=mc)
        arrays = {'cfg.ndc': numpy.array(int(ndc)), 'cfg.resolution': numpy.array(mc['resolution']),
                  'cfg.intrinsic': numpy.array(mc['intrinsic']), 'cfg.average_pose': numpy.array(mc['average_pose']),
                  'cfg.translation_scale': numpy.array(mc['translation_scale']),
                  'cfg.near': numpy.array(mc['near']), 'cfg.far': numpy.array(mc['far']),
                  'cfg.near_ndc': numpy.array(mc.get('near_ndc', 0.0)), 'cfg.far_ndc': numpy.array(mc.get('far_ndc', 1.0))}
        # case A: plain frame; case B: separate view pose + two secondary cameras
        pose, view_pose, sec = poses[3], poses[40 % len(poses)], [poses[10 % len(poses)], poses[77 % len(poses)]]
        arrays.update({'pose.render': pose, 'pose.view': view_pose, 'pose.secondary': numpy.stack(sec)})
        a = dp.create_test_data(pose, None, None, True)
        b = dp.create_test_data(pose, view_pose, sec, True)
        arrays.update({f'a.{k}': v.numpy() for k, v in a.items()})
        arrays.update({f'b.{k}': v.numpy() for k, v in b.items()})
        # retrieve_inference_outputs on synthetic outputs: out-of-range colours, exact .5 roundings, negative depths
        R = hw[0] * hw[1]
        rgb = g.uniform(-0.2, 1.2, size=(R, 3)).astype(numpy.float32)
        rgb[:256, 0] = (numpy.arange(256, dtype=numpy.float32) + 0.5) / 255.0      # round-half-to-even cases
        outs = {'rgb_fine': rgb, 'depth_fine': g.normal(2.0, 2.0, R).astype(numpy.float32),
                'depth_var_fine': g.normal(0.1, 0.2, R).astype(numpy.float32),
                'visibility2_fine': g.uniform(0, 1, size=(R, 2)).astype(numpy.float32)}
        if ndc:
            outs['depth_ndc_fine'] = g.uniform(-0.1, 1.0, R).astype(numpy.float32)
            outs['depth_var_ndc_fine'] = g.normal(0.0, 0.1, R).astype(numpy.float32)
        ret = dp.retrieve_inference_outputs({k: torch.from_numpy(v) for k, v in outs.items()})
        arrays.update({f'net.{k}': v for k, v in outs.items()})
        arrays.update({f'ret.{k}': v for k, v in ret.items()})
        path = os.path.join(GOLDEN, f'frame_{name}.npz')
        numpy.savez_compressed(path, **arrays)
        print(f'frame_{name}.npz: {os.path.getsize(path) / 1024:.0f} KiB, {len(arrays)} arrays, ndc={ndc}')


if __name__ == '__main__':
    main()
