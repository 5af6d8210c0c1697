"""Shared helpers for the parity tests (test infrastructure)."""
import json
import os

import numpy
import torch

from oracle import vipnerf_oracle as O

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), 'golden')


def load_npz(name):
    with numpy.load(os.path.join(GOLDEN, name)) as f:
        return {k: torch.from_numpy(f[k]) for k in f.files}


def load_npz_raw(name):
    """numpy arrays (no torch conversion): frame fixtures hold uint8 / 0-d arrays"""
    with numpy.load(os.path.join(GOLDEN, name)) as f:
        return {k: f[k] for k in f.files}


def split_io(arrays):
    inputs = {k[3:]: v for k, v in arrays.items() if k.startswith('in.')}
    outputs = {k[4:]: v for k, v in arrays.items() if k.startswith('out.')}
    return inputs, outputs


def manifest():
    with open(os.path.join(GOLDEN, 'MANIFEST.json')) as f:
        return json.load(f)


def rel_err(a: torch.Tensor, b: torch.Tensor):
    """max|a-b| / max|b| (the normalisation SURVEY.md section 8c uses) and the median of |a-b| / max|b|."""
    a, b = a.detach().double().cpu(), b.detach().double().cpu()
    scale = b.abs().max().clamp_min(1e-30)
    d = (a - b).abs() / scale
    return d.max().item(), d.median().item()


def to_cuda(batch):
    return {k: (v.cuda() if isinstance(v, torch.Tensor) else v) for k, v in batch.items()}


def state_dict_cuda(seed=0):
    return {k: v.cuda() for k, v in O.synth_state_dict(seed).items()}


# --------------------------------------------------------------------------------------
# Training step: the reference's four losses restated for the tests
# --------------------------------------------------------------------------------------

def training_loss(out, sup, weights=(1.0, 0.1, 0.001, 0.1)):
    """TotalLoss of the shipped training configs on a model output dict.  Reference: LossComputer01.py:33-51 with
    MSE01.py:25-66 (rgb_coarse + rgb_fine on indices_mask_nerf), VisibilityLoss01.py:26-68 (mutual-detach MAE between
    raw_visibility[..., 0] and the transmittance `visibility`), VisibilityPriorLoss01.py:25-88 (masked 1 - visibility2)
    and SparseDepthMSE01.py:27-70 (depth_fine on indices_mask_sparse_depth); weights of runs/training/train0012 at
    iteration >= 30000."""
    m_nerf = sup['indices_mask_nerf'].bool()
    m_depth = sup['indices_mask_sparse_depth'].bool()
    target = sup['target_rgb']
    mse = sum(torch.mean(torch.mean(torch.square(out[f'rgb_{t}'][m_nerf] - target[m_nerf]), dim=1))
              for t in ('coarse', 'fine'))

    def mae(a, b):
        return torch.mean(torch.abs(a - b))

    vis = 0
    for t in ('coarse', 'fine'):
        pred, tgt = out[f'raw_visibility_{t}'][..., 0], out[f'visibility_{t}']
        vis = vis + mae(pred, tgt.detach()) + mae(pred.detach(), tgt)
    prior_mask = sup['visibility_prior_masks'][m_nerf]
    prior = sum(torch.mean(torch.sum(prior_mask * (1 - out[f'visibility2_{t}'][m_nerf]), dim=1))
                for t in ('coarse', 'fine'))
    depth = torch.mean(torch.square(out['depth_fine'][m_depth] - sup['sparse_depth_values'][:, 0][m_depth]))
    parts = {'MSE01': mse, 'VisibilityLoss01': vis, 'VisibilityPriorLoss01': prior, 'SparseDepthMSE01': depth}
    total = weights[0] * mse + weights[1] * vis + weights[2] * prior + weights[3] * depth
    return total, parts


def split_train_golden(arrays):
    rays = {k[3:]: v for k, v in arrays.items() if k.startswith('in.')}
    sup = {k[4:]: v for k, v in arrays.items() if k.startswith('sup.')}
    draws = {k[5:]: v for k, v in arrays.items() if k.startswith('draw.')}
    outs = {k[4:]: v for k, v in arrays.items() if k.startswith('out.')}
    grads = {}
    for k, v in arrays.items():
        if k.startswith('grad.'):
            name, kind = k[5:].rsplit('.', 1)
            grads.setdefault(name, {})[kind] = v
    return rays, sup, draws, outs, grads


def check_grad_fingerprint(name, g, fp, rtol, report=None):
    """Compares a gradient tensor with the fingerprint oracle/make_golden_train.py stored (full tensor for the small
    parameters; sum / L2 norm / 256 strided samples for the large ones).  Errors are relative to the tensor's
    largest |gradient|."""
    g = g.detach().cpu()
    if 'full' in fp:
        ref = fp['full']
        scale = ref.abs().max().clamp_min(1e-30).item()
        err = ((g - ref).abs().max() / scale).item()
    else:
        flat = g.reshape(-1).double()
        idx = torch.linspace(0, flat.numel() - 1, 256).long()
        scale = float(fp['absmax'])
        err = ((flat[idx] - fp['samples'].double()).abs().max() / scale).item()
        err = max(err, abs(float(flat.norm()) - float(fp['norm'])) / float(fp['norm']))
    if report is not None:
        report[name] = err
    assert err <= rtol, f'{name}: gradient differs from the reference by {err:.3e} (> {rtol:.1e}) of its max'
