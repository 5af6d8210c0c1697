"""Host-side logic of the multi-GPU path on CPU: world_size-2 gloo process group, stub renderer.
Checks the partition, the single-gather protocol and that shard -> render -> gather reproduces the
un-sharded result bit-for-bit (the path is per-ray, so it must)."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from vipnerf_b200 import sharding


def test_shard_range_covers_everything():
    for n in (0, 1, 2, 7, 4096, 190512):
        for world in (1, 2, 3, 8):
            spans = [sharding.shard_range(n, r, world) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            for (a, b), (c, d) in zip(spans[:-1], spans[1:]):
                assert b == c and a <= b
            assert max(b - a for a, b in spans) <= (n + world - 1) // world


def _stub_render(batch):
    """A deterministic per-ray function with the output structure of the real renderer."""
    o, d = batch['rays_o'], batch['rays_d']
    # only exactly-rounded elementwise ops: CPU vectorised transcendentals differ by an ulp between batch sizes
    rgb = o * 3 + d
    depth = o[:, 0] * d[:, 0] + o[:, 1] * d[:, 1]
    alpha = o[:, :1] + torch.arange(5, dtype=torch.float32)[None, :] * d[:, 1:2]
    return {'rgb_fine': rgb, 'depth_fine': depth, 'alpha_fine': alpha}


def _worker(rank, world, port, n_rays, result_queue):
    os.environ['MASTER_ADDR'] = '127.0.0.1'
    os.environ['MASTER_PORT'] = str(port)
    dist.init_process_group('gloo', rank=rank, world_size=world)
    try:
        g = torch.Generator().manual_seed(0)
        batch = {'rays_o': torch.randn(n_rays, 3, generator=g), 'rays_d': torch.randn(n_rays, 3, generator=g),
                 'num_frames': 3}
        out = sharding.render_sharded(_stub_render, batch)
        if rank == 0:
            full = _stub_render(batch)
            ok = set(out) == set(full) and all(torch.equal(out[k], full[k]) for k in full)
            result_queue.put(bool(ok))
        else:
            assert out is None
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize('n_rays', [1, 5, 64, 1001])
def test_sharded_render_equals_unsharded_gloo(n_rays):
    with socket.socket() as s:
        s.bind(('127.0.0.1', 0))
        port = s.getsockname()[1]
    ctx = mp.get_context('spawn')
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, n_rays, q)) for r in range(2)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(timeout=120)
        assert p.exitcode == 0
    assert q.get(timeout=10) is True


def _grad_worker(rank, world, port, result_queue):
    os.environ['MASTER_ADDR'] = '127.0.0.1'
    os.environ['MASTER_PORT'] = str(port)
    dist.init_process_group('gloo', rank=rank, world_size=world)
    try:
        torch.manual_seed(0)
        net = torch.nn.Sequential(torch.nn.Linear(6, 16), torch.nn.ReLU(), torch.nn.Linear(16, 3))
        g = torch.Generator().manual_seed(1)
        x, y = torch.randn(10, 6, generator=g), torch.randn(10, 3, generator=g)
        lo, hi = sharding.shard_range(10, rank, world)
        # global-mean loss: every rank contributes its shard's share, gradients are summed
        loss = torch.sum(torch.square(net(x[lo:hi]) - y[lo:hi])) / x.shape[0]
        loss.backward()
        sharding.allreduce_gradients(net)
        if rank == 0:
            ref = torch.nn.Sequential(torch.nn.Linear(6, 16), torch.nn.ReLU(), torch.nn.Linear(16, 3))
            ref.load_state_dict(net.state_dict())
            torch.sum(torch.square(ref(x) - y) / x.shape[0]).backward()
            ok = all(torch.allclose(a.grad, b.grad, rtol=1e-5, atol=1e-7) for a, b in zip(net.parameters(), ref.parameters()))
            result_queue.put(bool(ok))
    finally:
        dist.destroy_process_group()


def test_allreduce_gradients_equals_full_batch_gloo():
    """Data-parallel training step: shard the rays, backward per rank, one all-reduce of the flat gradient buffer ==
    the gradients of the un-sharded batch."""
    with socket.socket() as s:
        s.bind(('127.0.0.1', 0))
        port = s.getsockname()[1]
    ctx = mp.get_context('spawn')
    q = ctx.Queue()
    procs = [ctx.Process(target=_grad_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(timeout=120)
        assert p.exitcode == 0
    assert q.get(timeout=10) is True
