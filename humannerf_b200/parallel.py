"""One process per GPU: ray sharding, render gather and the per-step gradient all-reduce
(SURVEY.md section 8e).  The reference's only parallelism is single-process nn.DataParallel
around the two MLPs (network.py:68-72,115-119); here rays are split contiguously over ranks,
each rank runs the whole per-ray path on its slice, and

* rendering needs one all_gather of the per-ray results (rgb, alpha, depth: 20 B / ray);
* training needs an all_gather of the rendered rgb (so every rank evaluates the identical
  patch loss) and ONE bucketed all-reduce of the gradients per step.

The all-reduce is cut at the replicated-input boundary ("cut B"): the summed tensors are the MLP
parameter gradients (595 207 floats), d(motion-weight volume) (819 200 floats) and
d(motion_scale_Rs, motion_Ts) (288 floats) - 5.7 MB instead of the 257.7 MB of all 64.4 M
parameter gradients ("cut A", 63.6 M of which belong to the volume decoder).  Every rank then
back-propagates the summed volume / pose gradients through its identical decoder, FK and
pose-decoder graphs, which gives the same parameter gradients as cut A up to fp32 summation order.
"""
from __future__ import annotations

import os
from typing import Dict, List, Sequence, Tuple

import torch
import torch.distributed as dist


def shard_bounds(n: int, rank: int, world: int) -> Tuple[int, int]:
    """Contiguous block [begin, end) of `n` rays owned by `rank` (sizes differ by at most one)."""
    base, rem = divmod(n, world)
    begin = rank * base + min(rank, rem)
    return begin, begin + base + (1 if rank < rem else 0)


def shard_sizes(n: int, world: int) -> List[int]:
    return [shard_bounds(n, r, world)[1] - shard_bounds(n, r, world)[0] for r in range(world)]


def all_gather_rows(local: torch.Tensor, n_total: int, group=None) -> torch.Tensor:
    """Concatenate per-rank row blocks (uneven sizes allowed) in rank order."""
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1:
        return local
    world = dist.get_world_size(group)
    sizes = shard_sizes(n_total, world)
    pad = max(sizes)
    buf = local.new_zeros((pad,) + tuple(local.shape[1:]))
    buf[:local.shape[0]] = local
    out = [torch.empty_like(buf) for _ in range(world)]
    dist.all_gather(out, buf, group=group)
    return torch.cat([o[:s] for o, s in zip(out, sizes)], dim=0)


class GatherRowsWithGrad(torch.autograd.Function):
    """all_gather whose backward returns this rank's slice of the upstream gradient.  Every rank
    computes the identical loss on the gathered rgb, so no reduction of d(rgb) is needed."""

    @staticmethod
    def forward(ctx, local, n_total):
        ctx.n_total = n_total
        return all_gather_rows(local, n_total)

    @staticmethod
    def backward(ctx, g):
        if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
            return g, None
        a, b = shard_bounds(ctx.n_total, dist.get_rank(), dist.get_world_size())
        return g[a:b].contiguous(), None


def flatten_bucket(tensors: Sequence[torch.Tensor]) -> torch.Tensor:
    return torch.cat([t.reshape(-1) for t in tensors]) if tensors else torch.empty(0)


def unflatten_bucket(flat: torch.Tensor, like: Sequence[torch.Tensor]) -> List[torch.Tensor]:
    out, o = [], 0
    for t in like:
        n = t.numel()
        out.append(flat[o:o + n].view_as(t))
        o += n
    return out


def allreduce_bucket(tensors: Sequence[torch.Tensor], group=None) -> None:
    """Sum `tensors` over ranks with one flat all-reduce (in place)."""
    if not tensors or not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1:
        return
    flat = flatten_bucket(tensors)
    dist.all_reduce(flat, op=dist.ReduceOp.SUM, group=group)
    for t, v in zip(tensors, unflatten_bucket(flat, tensors)):
        t.copy_(v)


class CutBReducer:
    """Gradient exchange at the replicated-input boundary.

    Usage inside a train step (see bench.py / tests/test_parallel_cpu.py):
        red = CutBReducer(network)
        with red.capture(frame_tensors):      # per-frame tensors the sharded path consumes
            loss.backward(inputs=red.leaf_inputs())   # -> grads of MLP params + boundary tensors
        red.allreduce()                        # one bucket
        red.finish_backward()                  # local backward through decoder / FK / pose decoder

    Precondition: every rank renders rays of the SAME frame in a step (same `dst_posevec`, priors, parameters), as
    SURVEY 8e's split of one step over ranks does.  Summing boundary gradients and back-propagating the sum through
    this rank's own front end is only the true gradient then.  `check_same_frame` asserts it (one tiny all-reduce;
    enabled by HN_CHECK_SAME_FRAME=1 or the `check` argument) - with per-rank frames use an all-reduce of the
    parameter gradients instead (cut A).
    """

    MLP_PREFIXES = ('non_rigid_mlp.', 'cnl_mlp.')

    def __init__(self, network: torch.nn.Module):
        self.network = network
        self.mlp_params = [p for n, p in network.named_parameters() if n.startswith(self.MLP_PREFIXES)]
        self.boundary: Dict[str, torch.Tensor] = {}

    def set_boundary(self, **tensors: torch.Tensor) -> None:
        """Register the per-frame tensors (outputs of the replicated front end: motion-weight volume,
        motion_scale_Rs, motion_Ts) whose gradients are exchanged instead of their producers'."""
        self.boundary = dict(tensors)

    def check_same_frame(self, group=None) -> None:
        """Raise unless the boundary tensors (functions of the frame and the replicated parameters only) are
        identical on every rank: max over ranks of a checksum equals its min."""
        if not self.boundary or not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1:
            return
        sums = torch.stack([t.detach().double().sum() for t in self.boundary.values()])
        pair = torch.stack([sums, -sums])
        dist.all_reduce(pair, op=dist.ReduceOp.MAX, group=group)
        if not torch.equal(pair[0], -pair[1]):
            raise RuntimeError('CutBReducer: ranks hold different frames / front-end outputs in this step; the cut-B '
                               'gradient exchange needs the same frame on every rank (use cut A otherwise)')

    def leaf_inputs(self) -> List[torch.Tensor]:
        return self.mlp_params + list(self.boundary.values())

    def exchange_and_finish(self, loss: torch.Tensor, group=None) -> int:
        """Backward to the boundary, one bucketed all-reduce, then the replicated tail.
        Returns the number of floats exchanged."""
        if os.environ.get('HN_CHECK_SAME_FRAME', '0') == '1' and not torch.cuda.is_current_stream_capturing():
            self.check_same_frame(group)
        leaves = self.leaf_inputs()
        grads = torch.autograd.grad(loss, leaves, retain_graph=True, allow_unused=True)
        grads = [torch.zeros_like(t) if g is None else g.contiguous() for g, t in zip(grads, leaves)]
        allreduce_bucket(grads, group=group)
        n_mlp = len(self.mlp_params)
        for p, g in zip(self.mlp_params, grads[:n_mlp]):
            if p.grad is None:
                p.grad = g
            else:
                p.grad.add_(g)          # in place: keeps the flat-buffer views of training.FusedAdam
        b_tensors = list(self.boundary.values())
        b_grads = grads[n_mlp:]
        live = [(t, g) for t, g in zip(b_tensors, b_grads) if t.requires_grad]
        if live:
            torch.autograd.backward([t for t, _ in live], [g for _, g in live])
        return int(sum(g.numel() for g in grads))
