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
