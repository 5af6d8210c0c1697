"""vipnerf_b200.hostio: one-buffer host I/O and the CUDA-graphed plugin call must deliver exactly what `model(batch)`
returns (same kernel, same arithmetic) - also after the pinned inputs are refilled with another batch."""
import pytest
import torch

from oracle import vipnerf_oracle as O

pytestmark = pytest.mark.gpu

KEYS = ('rgb_fine', 'depth_fine', 'depth_ndc_fine', 'acc_coarse', 'alpha_fine')


def _model(precision='bf16'):
    from vipnerf_b200.ModelFactory import get_model
    mlp = dict(num_samples=64, netdepth=8, netwidth=256, points_positional_encoding_degree=10,
               views_positional_encoding_degree=4, use_view_dirs=True, view_dependent_rgb=True, predict_visibility=True)
    cfg = {'data_loader': {'ndc': True},
           'model': dict(name='VipNeRFFused01', coarse_mlp=dict(mlp), fine_mlp=dict(mlp, num_samples=128), chunk=4096,
                         lindisp=False, netchunk=16384, perturb=True, raw_noise_std=1.0, white_bkgd=False,
                         precision=precision)}
    model = get_model(cfg, None)
    model.load_state_dict(O.synth_state_dict(0))
    return model.cuda().eval()


def test_graphed_render_equals_plugin_call(built_library):
    from vipnerf_b200 import hostio
    model = _model()
    a, b = O.make_rays('fern', 777, seed=2), O.make_rays('fern', 777, seed=9)
    g = hostio.GraphedRender(model, a, KEYS)
    with torch.no_grad():
        for batch in (a, b, a):
            got = g(batch)
            g.synchronize()
            ref = model({k: v.cuda() for k, v in batch.items()})
            for k in KEYS:
                assert torch.equal(got[k], ref[k].cpu()), k
    assert g.h2d_bytes >= sum(v.numel() * 4 for v in a.values()) and g.d2h_bytes >= 777 * (3 + 1 + 1 + 1 + 192) * 4
    # a weight update after the capture: the graph must not replay the stale packed weights
    with torch.no_grad():
        model.fine_model.pts_linears[3].weight.mul_(1.05)
        got = g(b)
        g.synchronize()
        ref = model({k: v.cuda() for k, v in b.items()})
        for k in KEYS:
            assert torch.equal(got[k], ref[k].cpu()), k


def test_out_tensors_are_written_in_place(built_library):
    """`model(batch, out={...})`: the named outputs are the caller's tensors (what PeerGather / FlatBuffers rely on),
    the other keys are allocated as usual; wrong shapes are rejected."""
    model = _model()
    batch = {k: v.cuda() for k, v in O.make_rays('fern', 130, seed=4).items()}
    mine = {'rgb_fine': torch.full((130, 3), -1.0, device='cuda'), 'depth_coarse': torch.full((130,), -1.0, device='cuda')}
    with torch.no_grad():
        ref = model(dict(batch))
        out = model(dict(batch), out=mine)
        assert out['rgb_fine'] is mine['rgb_fine'] and out['depth_coarse'] is mine['depth_coarse']
        assert set(out) == set(ref)
        for k in ref:
            assert torch.equal(out[k], ref[k]), k
        with pytest.raises(ValueError):
            model(dict(batch), out={'rgb_fine': torch.empty((129, 3), device='cuda')})
