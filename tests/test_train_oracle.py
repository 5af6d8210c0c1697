"""CPU: the oracle's training mode (forward with the reference's random draws + torch autograd as the gradient
oracle) against the golden training step of the unmodified reference (tests/golden/train_*.npz, made by
oracle/make_golden_train.py)."""
import pytest
import torch

from oracle import vipnerf_oracle as O
from tests import helpers as H


@pytest.mark.parametrize('scene', ['fern', 'dtu'])
def test_draw_order_matches_reference(scene):
    """Re-seeding and drawing through the oracle's mirror of the reference's RNG consumption order reproduces the
    stored draws; that they ARE the reference's draws is what test_oracle_training_step_matches_reference shows."""
    _, _, draws, _, _ = H.split_train_golden(H.load_npz(f'train_{scene}.npz'))
    torch.manual_seed(77)
    again = O.draw_training_randoms(48, 64, 128, chunk=32, netchunk=1000, perturb=True, raw_noise_std=1.0)
    assert set(again) == set(draws)
    for k in draws:
        assert torch.equal(again[k], draws[k]), k


@pytest.mark.parametrize('scene', ['fern', 'dtu'])
def test_oracle_training_step_matches_reference(scene):
    rays, sup, draws, outs, grads = H.split_train_golden(H.load_npz(f'train_{scene}.npz'))
    ndc = O.SCENES[scene]['ndc']
    sd = {k: v.clone().requires_grad_(True) for k, v in O.synth_state_dict(0).items()}
    out = O.render(sd, rays, ndc=ndc, chunk=32, netchunk=1000, train_randoms=draws)
    for k, ref in outs.items():
        mx, _ = H.rel_err(out[k], ref)
        assert mx <= 2e-5, f'{k}: {mx:.2e}'
    total, parts = H.training_loss(out, sup)
    total.backward()
    golden = H.load_npz(f'train_{scene}.npz')
    assert abs(float(total.detach()) - float(golden['loss.total'])) <= 1e-5 * abs(float(golden['loss.total']))
    for name, value in parts.items():
        assert abs(float(value) - float(golden[f'loss.{name}'])) <= 2e-5 * max(1.0, abs(float(golden[f'loss.{name}']))), name
    report = {}
    for name, fp in grads.items():
        H.check_grad_fingerprint(name, sd[name].grad, fp, 1e-5, report)   # measured 3e-7
    assert len(report) == 48


@pytest.mark.parametrize('n_rays,chunk,netchunk,perturb,noise,has_fine', [
    (48, 32, 1000, True, 1.0, True), (100, 4096, 16384, True, 1.0, True), (70, 64, None, False, 0.5, True),
    (33, 16, 500, True, 0.0, False)])
def test_plugin_draw_order_matches_oracle(n_rays, chunk, netchunk, perturb, noise, has_fine):
    """The product's own generator mirror (vipnerf_b200.training.draw_training_randoms, what the plugin calls in train
    mode) consumes torch's CPU generator exactly like the oracle's, which the golden training step pins to the
    reference: same keys, bit-identical tensors, same generator state afterwards."""
    from vipnerf_b200 import training
    torch.manual_seed(123)
    a = training.draw_training_randoms(n_rays, 64, 128, chunk, netchunk, perturb, noise, has_fine)
    tail_a = torch.rand(4)
    torch.manual_seed(123)
    b = O.draw_training_randoms(n_rays, 64, 128, chunk, netchunk if netchunk else 10 ** 9, perturb, noise, has_fine)
    tail_b = torch.rand(4)
    assert set(a) == set(b)
    for k in a:
        assert torch.equal(a[k], b[k]), k
    assert torch.equal(tail_a, tail_b)


def test_loss_computer_torch_path_matches_reference_losses():
    """LossComputerFused01's torch evaluation (used for validation batches / foreign output dicts) on the golden
    outputs of the reference's training step reproduces the loss values of the reference's own LossComputer01."""
    import pytest
    from tests import helpers as H
    from vipnerf_b200.LossComputerFused01 import LossComputer
    cfg = {'model': {'coarse_mlp': {}, 'fine_mlp': {}},
           'losses': [{'name': 'MSE01', 'weight': 1}, {'name': 'VisibilityLoss01', 'weight': 0.1},
                      {'name': 'VisibilityPriorLoss01', 'iter_weights': {'0': 0, '30000': 0.001}},
                      {'name': 'SparseDepthMSE01', 'weight': 0.1}]}
    for scene in ('fern', 'dtu'):
        arrays = H.load_npz(f'train_{scene}.npz')
        rays, sup, draws, outs, grads = H.split_train_golden(arrays)
        batch = dict(rays)
        batch.update(sup)
        losses = LossComputer(cfg).compute_losses(batch, outs)
        for name in ('MSE01', 'VisibilityLoss01', 'VisibilityPriorLoss01', 'SparseDepthMSE01'):
            ref = float(arrays[f'loss.{name}'])
            assert abs(float(losses[name]['loss_value']) - ref) <= 2e-5 * max(abs(ref), 1e-3), (scene, name)
        assert abs(float(losses['TotalLoss']) - float(arrays['loss.total'])) <= 2e-5 * abs(float(arrays['loss.total']))
    with pytest.raises(RuntimeError):
        LossComputer({'model': {}, 'losses': [{'name': 'SomethingElse01', 'weight': 1}]})
