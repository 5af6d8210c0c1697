"""Generates tests/golden/*.npz by running the UNMODIFIED reference (imported from /root/reference) on
seeded inputs.  Run in the build container only:  python oracle/make_golden.py
TEST INFRASTRUCTURE ONLY.  The fixtures are what pins oracle/vipnerf_oracle.py (and through it the CUDA
path) to the reference; the reference itself ships no golden vectors for this path.
"""
from __future__ import annotations

import hashlib
import json
import os
import sys

import numpy
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import ref_loader  # noqa: E402
from oracle import vipnerf_oracle as O  # noqa: E402

GOLDEN = os.path.join(ROOT, 'tests', 'golden')


def state_dict_digest(sd) -> str:
    h = hashlib.sha256()
    for k in sorted(sd):
        h.update(k.encode())
        h.update(sd[k].contiguous().numpy().tobytes())
    return h.hexdigest()


def save(name, **arrays):
    path = os.path.join(GOLDEN, name)
    numpy.savez_compressed(path, **{k: (v.detach().numpy() if isinstance(v, torch.Tensor) else numpy.asarray(v))
                                    for k, v in arrays.items()})
    print(f'{name}: {os.path.getsize(path) / 1024:.0f} KiB, {len(arrays)} arrays')


def main():
    os.makedirs(GOLDEN, exist_ok=True)
    torch.set_num_threads(max(1, os.cpu_count() or 1))
    ref_loader.import_reference()
    from models.VipNeRF01 import MLP, VipNeRF  # the reference classes

    sd = O.synth_state_dict(0)
    manifest = {
        'reference_commit': 'a7f1d81de376c30348663d153e40940c212474d5',
        'torch': torch.__version__,
        'state_dict_seed': 0,
        'state_dict_sha256': state_dict_digest(sd),
        'files': {},
    }

    # ---- whole-path fixtures: reference model.forward on 64 rays (retraw + secondary views) and on
    #      512 rays (eval keys only, per-ray maps) for the NDC (LLFF fern) and world-space (DTU) cameras
    for scene in ('fern', 'dtu'):
        ndc = O.SCENES[scene]['ndc']
        model = ref_loader.build_reference_model(sd, ndc)
        small = O.make_rays(scene, 64, seed=3, n_sec_views=2)
        with torch.no_grad():
            out = model(dict(small), retraw=True, sec_views_vis=True)
        arrays = {f'in.{k}': v for k, v in small.items()}
        arrays.update({f'out.{k}': v for k, v in out.items() if not k.startswith('raw_rgb_view_dependent')})
        save(f'render_{scene}_retraw64.npz', **arrays)
        big = O.make_rays(scene, 512, seed=5)
        with torch.no_grad():
            out = model(dict(big))
        arrays = {f'in.{k}': v for k, v in big.items()}
        arrays.update({f'out.{k}': v for k, v in out.items() if not k.startswith('alpha')})
        save(f'render_{scene}_eval512.npz', **arrays)
        manifest['files'][scene] = {'retraw': 'seed 3, 64 rays, V=2', 'eval': 'seed 5, 512 rays'}

    # ---- stage fixtures straight from the reference's functions
    g = torch.Generator().manual_seed(11)
    # sample_pdf (static method :229-262): random pdfs plus the degenerate rows (all-zero weights,
    # one-hot weights, a denom<1e-5 plateau)
    bins = torch.sort(torch.rand(48, 63, generator=g) * 4 + 0.5, dim=-1)[0]
    w = torch.rand(48, 62, generator=g) ** 4
    w[0] = 0
    w[1] = 0
    w[1, 30] = 1
    w[2, 5:40] = 0
    w[3] = 1e-9
    samples = VipNeRF.sample_pdf(bins, w, 128, det=True)
    u = torch.rand(48, 128, generator=g)
    torch.manual_seed(1234)
    samples_rand = VipNeRF.sample_pdf(bins, w, 128, det=False)   # consumes torch.rand(48,128) after seed 1234
    torch.manual_seed(1234)
    u_rand = torch.rand(48, 128)
    save('stage_sample_pdf.npz', bins=bins, weights=w, samples_det=samples, u_rand=u_rand, samples_rand=samples_rand)

    # convert_depth_from_ndc (static :386-403), including z_ndc == 1 exactly
    rays = O.make_rays('fern', 32, seed=2)
    z_ndc = torch.rand(32, 64, generator=g)
    z_ndc[:, -1] = 1.0
    save('stage_depth_from_ndc.npz', z_ndc=z_ndc, rays_o=rays['rays_o'], rays_d=rays['rays_d'],
         depth=VipNeRF.convert_depth_from_ndc(z_ndc, rays['rays_o'], rays['rays_d']))

    # positional encoders (:416-448 via MLP.get_positional_encoder :494-507)
    x = (torch.rand(64, 3, generator=g) * 2 - 1) * torch.tensor([1.5, 1.5, 5.0])
    enc10, dim10 = MLP.get_positional_encoder(10)
    enc4, dim4 = MLP.get_positional_encoder(4)
    save('stage_posenc.npz', x=x, enc10=enc10(x), enc4=enc4(x), dims=numpy.array([dim10, dim4]))

    # one MLP forward (:509-535) on 192 points with two secondary view directions per point
    cfg = ref_loader.reference_configs(True)
    mlp = MLP(cfg, cfg['model']['fine_mlp']).eval()
    mlp.load_state_dict(O.split_state_dict(sd, 'fine_model'))
    pts = (torch.rand(192, 3, generator=g) * 2 - 1) * torch.tensor([1.2, 1.2, 1.0])
    vd = torch.nn.functional.normalize(torch.randn(192, 3, generator=g), dim=-1)
    vd2 = torch.nn.functional.normalize(torch.randn(192, 2, 3, generator=g), dim=-1)
    with torch.no_grad():
        o = mlp({'pts': pts, 'view_dirs': vd, 'view_dirs2': vd2})
    save('stage_mlp.npz', pts=pts, view_dirs=vd, view_dirs2=vd2, sigma=o['sigma'], rgb=o['rgb'],
         visibility=o['visibility'], visibility2=o['visibility2'])

    with open(os.path.join(GOLDEN, 'MANIFEST.json'), 'w') as f:
        json.dump(manifest, f, indent=1)
    print('manifest', manifest['state_dict_sha256'][:16])


if __name__ == '__main__':
    main()
