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


_gather_buffers: Dict[tuple, Dict[str, torch.Tensor]] = {}


def _full_buffers(local: Dict[str, torch.Tensor], per: int, world: int) -> Dict[str, torch.Tensor]:
    """Destination arrays [world * per, ...] per key, allocated once per (keys, shapes, device) and reused."""
    sig = tuple(sorted((k, tuple(v.shape[1:]), str(v.dtype), str(v.device)) for k, v in local.items())) + (per, world)
    bufs = _gather_buffers.get(sig)
    if bufs is None:
        bufs = {k: torch.empty((world * per,) + tuple(v.shape[1:]), dtype=v.dtype, device=v.device) for k, v in local.items()}
        _gather_buffers[sig] = bufs
    return bufs


def gather_outputs(local: Dict[str, torch.Tensor], n_rays: int, group: Optional[dist.ProcessGroup] = None,
                   dst: int = 0) -> Optional[Dict[str, torch.Tensor]]:
    """The single collective of the path: the ranks' per-ray outputs land on `dst` (None elsewhere), every key in ONE
    grouped exchange (NCCL: one ncclGroup of sends / receives) straight into preallocated destination arrays - no packing
    before, no concatenation or copies after: rank r's rows of key k are received into rows [r*per, r*per + n_r) of the
    destination array of k, which is returned as is (a view of its first n_rays rows).  The returned tensors are reused
    by the next call with the same signature."""
    world = dist.get_world_size(group)
    rank = dist.get_rank(group)
    per = (n_rays + world - 1) // world
    keys = sorted(local)
    if world == 1:
        return {k: local[k] for k in keys}
    ops = []
    if rank == dst:
        full = _full_buffers(local, per, world)
        for r in range(world):
            lo, hi = shard_range(n_rays, r, world)
            if hi <= lo:
                continue
            for k in keys:
                if r == rank:
                    full[k][lo:hi].copy_(local[k])
                else:
                    ops.append(dist.P2POp(dist.irecv, full[k][lo:hi], dist.get_global_rank(group, r) if group is not None else r, group))
    else:
        lo, hi = shard_range(n_rays, rank, world)
        if hi > lo:
            for k in keys:
                ops.append(dist.P2POp(dist.isend, local[k].contiguous(), dist.get_global_rank(group, dst) if group is not None else dst, group))
    if ops:
        for work in dist.batch_isend_irecv(ops):
            work.wait()
    if rank == dst:
        return {k: full[k][:n_rays] for k in keys}
    return None


class PeerGather:
    """Fused compute + gather ("direct NVLink peer stores from the kernel epilogue
    into rank 0's image buffer"): the destination arrays live in torch symmetric memory (CUDA VMM allocations every
    rank of the node maps), every rank hands the render kernel OUTPUT POINTERS INTO `dst`'s arrays - its ray warps then
    store the finished per-ray maps over NVLink as they composite them - and one device-side barrier on the stream
    tells `dst` that all shards have landed.  No NCCL call, no staging copy, nothing to unpack: on `dst` the arrays are
    the frame.

        pg = PeerGather({'rgb_fine': (3,), 'depth_fine': ()}, n_rays, device)
        out = model(shard, out=pg.local_outputs())        # rendered straight into dst's memory
        frame = pg.finish()                                # barrier; dict of [n_rays, ...] views on dst, None elsewhere
    """

    def __init__(self, key_shapes: Dict[str, tuple], n_rays: int, device, group: Optional[dist.ProcessGroup] = None,
                 dst: int = 0):
        import torch.distributed._symmetric_memory as symm_mem
        group = group if group is not None else dist.group.WORLD
        self.world, self.rank, self.dst, self.n_rays = dist.get_world_size(group), dist.get_rank(group), dst, n_rays
        self.per = (n_rays + self.world - 1) // self.world
        self.keys = sorted(key_shapes)
        self.offsets, off = {}, 0
        for k in self.keys:
            width = 1
            for d in key_shapes[k]:
                width *= int(d)
            self.offsets[k] = (off, width, tuple(key_shapes[k]))
            off += ((self.world * self.per * width + 63) // 64) * 64       # 256-byte aligned regions
        self.total = max(off, 64)
        self.buf = symm_mem.empty(self.total, dtype=torch.float32, device=device)
        self.handle = symm_mem.rendezvous(self.buf, group)
        # dst's buffer as seen from this rank (on dst: the local buffer itself)
        self.dst_buf = self.buf if self.rank == dst else self.handle.get_buffer(dst, (self.total,), torch.float32)

    def view(self, base: torch.Tensor, k: str, lo: int, hi: int) -> torch.Tensor:
        """Rows [lo, hi) of key k inside a flat buffer with this object's layout (the symmetric buffer, or a host copy of it)."""
        off, width, shape = self.offsets[k]
        return base[off + lo * width: off + hi * width].view((hi - lo,) + shape)

    def local_outputs(self) -> Dict[str, torch.Tensor]:
        """This rank's rows of every key, as tensors that alias `dst`'s arrays (peer memory for rank != dst)."""
        lo, hi = shard_range(self.n_rays, self.rank, self.world)
        return {k: self.view(self.dst_buf, k, lo, hi) for k in self.keys}

    def finish(self) -> Optional[Dict[str, torch.Tensor]]:
        """Stream-ordered barrier over the ranks (symmetric-memory signal pads); afterwards `dst` owns the whole result."""
        self.handle.barrier(channel=0)
        if self.rank != self.dst:
            return None
        return {k: self.view(self.buf, k, 0, self.n_rays) for k in self.keys}

    def release(self) -> None:
        """Second barrier: nobody may start overwriting dst's arrays (next frame) before dst has consumed them."""
        self.handle.barrier(channel=1)


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
