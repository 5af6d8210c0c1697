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
