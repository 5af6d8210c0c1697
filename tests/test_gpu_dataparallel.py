"""The plugin under torch.nn.DataParallel with a device LIST, the way the reference wraps its model
(Tester01.py:39-49, Trainer01.py:517-519 with configs['device'] = [0, 1], NerfLlffTrainerTester01.py:329):
replicate() builds replicas without nn.Parameters, one Python thread per GPU calls forward concurrently, the outputs
are gathered on device 0 and the backward reduces the replicas' gradients onto the master parameters."""
import copy

import pytest
import torch

from oracle import vipnerf_oracle as O

pytestmark = pytest.mark.gpu


def _configs(precision):
    mlp = dict(num_samples=64, netdepth=8, netwidth=256, points_positional_encoding_degree=10,
               views_positional_encoding_degree=4, use_view_dirs=True, view_dependent_rgb=True, predict_visibility=True)
    return {'data_loader': {'ndc': True},
            'model': dict(name='VipNeRFFused01', coarse_mlp=dict(mlp), fine_mlp=dict(mlp, num_samples=128), chunk=4096,
                          lindisp=False, netchunk=16384, perturb=False, raw_noise_std=0.0, white_bkgd=False,
                          precision=precision)}


def _loss(out):
    return (out['rgb_fine'].square().mean() + out['rgb_coarse'].mean() + 0.1 * out['depth_fine'].mean()
            + 0.01 * out['visibility2_fine'].mean())


def test_replica_tensors_are_read_by_attribute(built_library):
    """Runs on one device: a hand-made replica (what DataParallel.replicate produces) has no parameters() yet must
    expose its 24 tensors; the packed-weight cache must not be used for replicas."""
    from vipnerf_b200.ModelFactory import get_model
    model = get_model(_configs('bf16'), None)
    model.load_state_dict(O.synth_state_dict(0))
    model = model.cuda().eval()
    replicas = torch.nn.parallel.replicate(model, [0])
    rep = replicas[0]
    assert len(list(rep.coarse_model.pts_linears[0].parameters())) == 0
    assert len(rep.coarse_model.named_tensors()) == 24
    batch = {k: v.cuda() for k, v in O.make_rays('fern', 300, seed=3).items()}
    with torch.no_grad():
        a = model(dict(batch))
        b = rep(dict(batch))
    for k in a:
        assert torch.equal(a[k], b[k]), k


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason='needs two CUDA devices')
def test_dataparallel_two_devices_render_step_render(built_library):
    from vipnerf_b200.ModelFactory import get_model
    single = get_model(_configs('bf16'), None)
    single.load_state_dict(O.synth_state_dict(0))
    single = single.to('cuda:0')
    wrapped = torch.nn.DataParallel(copy.deepcopy(single), device_ids=[0, 1])
    batch = {k: v.to('cuda:0') for k, v in O.make_rays('fern', 1001, seed=3, n_sec_views=1).items()}

    def render(m):
        m.eval()
        with torch.no_grad():
            return m(dict(batch), retraw=False, sec_views_vis=True)

    a, b = render(single), render(wrapped)
    assert set(a) == set(b)
    for k in a:
        assert b[k].device.index == 0 and torch.equal(a[k], b[k]), k

    # one optimizer step through each (Trainer01.py:93-102), then both must render the same again: the replicas'
    # packed weights may not be stale
    grads = {}
    for name, m in (('single', single), ('wrapped', wrapped)):
        m.train()
        opt = torch.optim.Adam(m.parameters(), lr=5e-4)
        opt.zero_grad()
        _loss(m(dict(batch))).backward()
        grads[name] = {k.replace('module.', ''): p.grad.clone() for k, p in m.named_parameters()}
        opt.step()
    for k, g in grads['single'].items():
        err = ((grads['wrapped'][k] - g).abs().max() / g.abs().max().clamp_min(1e-30)).item()
        assert err <= 2e-3, (k, err)       # same arithmetic, the ray sum split in two halves
    a2, b2 = render(single), render(wrapped)
    assert not torch.equal(a2['rgb_fine'], a['rgb_fine'])      # the step moved the weights
    for k in ('rgb_fine', 'depth_fine', 'visibility2_fine'):
        d = ((a2[k] - b2[k]).abs().max() / a2[k].abs().max()).item()
        assert d <= 2e-2, (k, d)          # weights after one Adam step agree to the gradient tolerance above
