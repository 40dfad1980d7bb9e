"""Multi-rank host logic on CPU: ray sharding, uneven all_gather with gradient, and the cut-B
gradient exchange (one bucketed all-reduce at the replicated-input boundary), world_size 2, gloo."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from humannerf_b200.parallel import (CutBReducer, GatherRowsWithGrad, all_gather_rows, allreduce_bucket,
                                     flatten_bucket, shard_bounds, shard_sizes, unflatten_bucket)


def test_shard_bounds_partition():
    for n in (0, 1, 7, 6144, 196522):
        for world in (1, 2, 3, 8):
            spans = [shard_bounds(n, r, world) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(a[1] == b[0] for a, b in zip(spans[:-1], spans[1:]))
            sizes = shard_sizes(n, world)
            assert sum(sizes) == n and max(sizes) - min(sizes) <= 1


def test_bucket_roundtrip():
    ts = [torch.randn(3, 4), torch.randn(5), torch.randn(2, 2, 2)]
    flat = flatten_bucket(ts)
    back = unflatten_bucket(flat, ts)
    assert all(torch.equal(a, b) for a, b in zip(ts, back))
    allreduce_bucket(ts)          # no process group: no-op


class Toy(torch.nn.Module):
    """decoder (replicated, heavy) -> boundary tensor -> per-ray MLP (sharded rays)."""

    def __init__(self):
        super().__init__()
        self.mweight_vol_decoder = torch.nn.Linear(4, 6)
        self.cnl_mlp = torch.nn.Linear(6, 3)
        self.boundary_tensors = {}

    def forward(self, rays, code):
        vol = torch.tanh(self.mweight_vol_decoder(code))          # per-frame, identical on all ranks
        self.boundary_tensors = {'vol': vol}
        return torch.sigmoid(self.cnl_mlp(rays * vol))


def _free_port():
    s = socket.socket()
    s.bind(('127.0.0.1', 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, ret):
    os.environ.update(MASTER_ADDR='127.0.0.1', MASTER_PORT=str(port))
    dist.init_process_group('gloo', rank=rank, world_size=world)
    try:
        torch.manual_seed(0)
        net = Toy()
        rays = torch.randn(11, 6)
        code = torch.randn(4)
        target = torch.randn(11, 3)
        a, b = shard_bounds(11, rank, world)
        # strong-scaling step: local rays -> gather rgb -> identical loss everywhere -> cut-B exchange
        out = net(rays[a:b], code)
        rgb = GatherRowsWithGrad.apply(out, 11)
        assert rgb.shape == (11, 3)
        loss = ((rgb - target) ** 2).mean()
        red = CutBReducer(net)
        red.set_boundary(**net.boundary_tensors)
        n = red.exchange_and_finish(loss)
        assert n == 6 * 3 + 3 + 6
        grads = {k: p.grad.clone() for k, p in net.named_parameters()}
        # single-process reference
        ref = Toy()
        ref.load_state_dict(net.state_dict())
        ((ref(rays, code) - target) ** 2).mean().backward()
        for k, p in ref.named_parameters():
            assert torch.allclose(grads[k], p.grad, atol=1e-6), k
        g = all_gather_rows(out.detach(), 11)
        assert torch.allclose(g, ref(rays, code).detach(), atol=1e-6)
        ret[rank] = True
    finally:
        dist.destroy_process_group()


def test_cut_b_exchange_world2():
    port = _free_port()
    mgr = mp.Manager()
    ret = mgr.dict()
    mp.spawn(_worker, args=(2, port, ret), nprocs=2, join=True)
    assert ret.get(0) and ret.get(1)
