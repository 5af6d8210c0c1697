"""Multi-GPU rendering: rays shard embarrassingly (every stage of the path is per-ray), so each rank renders a
contiguous ray range with its own copy of the packed weights and ONE gather brings the per-ray outputs to
rank 0.  This replaces the reference's single-process torch.nn.DataParallel scatter / replicate / gather
(src/Tester01.py:42, src/Trainer01.py:517) with one process per GPU over
torch.distributed (NCCL on GPUs; the same code runs over gloo for the CPU tests).

Partition: rank r of G gets rays [r*ceil(N/G), min(N, (r+1)*ceil(N/G))) - what DataParallel.scatter does along
dim 0.  No other collective exists on the render path.  Training (row f1) has one exchange step of its own: the
parameter gradients of the ranks' ray shards are summed (`allreduce_gradients`), which is what DataParallel's
backward does when it reduces the replicas' gradients onto the master copy (src/Trainer01.py:517).
"""
from __future__ import annotations

from typing import Callable, Dict, Optional, Tuple

import torch
import torch.distributed as dist

RAY_KEYS = ('rays_o', 'rays_d', 'view_dirs', 'near', 'far', 'rays_o_ndc', 'rays_d_ndc', 'near_ndc', 'far_ndc',
            'rays_o2', 'pixel_id')


def shard_range(n_rays: int, rank: int, world_size: int) -> Tuple[int, int]:
    per = (n_rays + world_size - 1) // world_size
    lo = min(n_rays, rank * per)
    return lo, min(n_rays, lo + per)


def shard_batch(batch: Dict[str, object], rank: int, world_size: int) -> Dict[str, object]:
    """Slices every per-ray tensor of a reference-style input dict to this rank's ray range; everything else
    (common_data, num_frames, ...) passes through, like VipNeRF.batchify_rays does per chunk (VipNeRF01.py:56-62)."""
    n = batch['rays_o'].shape[0]
    lo, hi = shard_range(n, rank, world_size)
    out = {}
    for k, v in batch.items():
        if isinstance(v, torch.Tensor) and v.dim() > 0 and v.shape[0] == n and k in RAY_KEYS:
            out[k] = v[lo:hi]
        else:
            out[k] = v
    return out


def gather_outputs(local: Dict[str, torch.Tensor], n_rays: int, group: Optional[dist.ProcessGroup] = None,
                   dst: int = 0) -> Optional[Dict[str, torch.Tensor]]:
    """The single collective of the path: concatenates the ranks' per-ray outputs on `dst` (None elsewhere).
    All keys are packed into one [rays, width] buffer so exactly one gather is issued regardless of how many
    maps were rendered; short last shards are padded to the common shard size."""
    world = dist.get_world_size(group)
    rank = dist.get_rank(group)
    per = (n_rays + world - 1) // world
    keys = sorted(local)
    widths = []
    for k in keys:
        w = 1
        for dim in local[k].shape[1:]:
            w *= int(dim)
        widths.append(w)
    first = local[keys[0]]
    packed = torch.zeros((per, sum(widths)), dtype=torch.float32, device=first.device)
    n_local = first.shape[0]
    col = 0
    for k, w in zip(keys, widths):
        packed[:n_local, col:col + w] = local[k].reshape(n_local, w)
        col += w
    if rank == dst:
        parts = [torch.empty_like(packed) for _ in range(world)]
        dist.gather(packed, parts, dst=dst, group=group)
        full = torch.cat(parts, dim=0)[:n_rays]
        out, col = {}, 0
        for k, w in zip(keys, widths):
            out[k] = full[:, col:col + w].reshape((n_rays,) + tuple(local[k].shape[1:])).contiguous()
            col += w
        return out
    dist.gather(packed, None, dst=dst, group=group)
    return None


def render_sharded(render_fn: Callable[[Dict[str, object]], Dict[str, torch.Tensor]], batch: Dict[str, object],
                   group: Optional[dist.ProcessGroup] = None, dst: int = 0) -> Optional[Dict[str, torch.Tensor]]:
    """Renders `batch` (the FULL ray set, present on every rank) across the ranks of `group`: this rank renders
    its shard with `render_fn` (e.g. a VipNeRFFused module) and the results are gathered on `dst`."""
    world = dist.get_world_size(group)
    rank = dist.get_rank(group)
    n = batch['rays_o'].shape[0]
    local = render_fn(shard_batch(batch, rank, world))
    local = {k: v for k, v in local.items() if isinstance(v, torch.Tensor)}
    return gather_outputs(local, n, group, dst)


def allreduce_gradients(module: torch.nn.Module, group: Optional[dist.ProcessGroup] = None, average: bool = False) -> None:
    """The one collective of a data-parallel training step: sums (or averages) `.grad` of every parameter over the
    ranks with ONE all-reduce of a flat buffer (2.4 MB of fp32 per MLP pair), then scatters the result back into the
    gradient tensors.  Call between `loss.backward()` and `optimizer.step()`.  A loss that is a mean over the GLOBAL
    batch must be scaled by the local share before backward (sum semantics), or use average=True for per-rank means."""
    params = [p for p in module.parameters() if p.grad is not None]
    if not params:
        return
    flat = torch.cat([p.grad.reshape(-1) for p in params])
    dist.all_reduce(flat, op=dist.ReduceOp.SUM, group=group)
    if average:
        flat /= dist.get_world_size(group)
    off = 0
    for p in params:
        n = p.numel()
        p.grad.copy_(flat[off:off + n].view_as(p.grad))
        off += n
