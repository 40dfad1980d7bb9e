"""One process per GPU: ray sharding, render gather and the per-step gradient all-reduce
(SURVEY.md section 8e).  The reference's only parallelism is single-process nn.DataParallel
around the two MLPs (network.py:68-72,115-119); here rays are split contiguously over ranks,
each rank runs the whole per-ray path on its slice, and

* rendering needs one all_gather of the per-ray results (rgb, alpha, depth: 20 B / ray);
* training needs an all_gather of the rendered rgb (so every rank evaluates the identical
  patch loss) and ONE bucketed all-reduce of the gradients per step.

Two ways to split a training step (bench.py --scaling):
  strong  ONE step's rays (<= n_patches * patch^2) are split contiguously over the ranks: forward on the
          local slice, all_gather of the rendered rgb (GatherRowsWithGrad: 12 B / ray), the identical patch loss on
          every rank, backward into the local rows only, gradients SUMMED over ranks;
  weak    every rank renders its own patches of the same frame; loss / world, gradients summed = the mean.

The all-reduce is cut at the replicated-input boundary ("cut B"): the summed tensors are the MLP
parameter gradients (595 207 floats), d(motion-weight volume) (819 200 floats) and
d(motion_scale_Rs, motion_Ts) (288 floats) - 5.7 MB instead of the 257.7 MB of all 64.4 M
parameter gradients ("cut A", 63.6 M of which belong to the volume decoder).  Every rank then
back-propagates the summed volume / pose gradients through its identical decoder, FK and
pose-decoder graphs, which gives the same parameter gradients as cut A up to fp32 summation order.
The exchange is two buckets: the boundary gradients (3.3 MB) are on the critical path of the replicated tail; the
MLP parameter gradients (2.4 MB) are only consumed by the optimizer, so their all-reduce is issued asynchronously and
waited for after the tail's backward - it runs under the decoder backward.
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
    if local.is_cuda and dist.get_backend(group) == 'gloo':
        # gloo moves CUDA tensors only through all_reduce / broadcast (single-GPU test rigs): every rank adds its
        # rows into a zero buffer - same result as the gather
        out = local.new_zeros((n_total,) + tuple(local.shape[1:]))
        a = sum(sizes[:dist.get_rank(group)])
        out[a:a + local.shape[0]] = local
        dist.all_reduce(out, op=dist.ReduceOp.SUM, group=group)
        return out
    if min(sizes) == max(sizes):
        # equal slices: one collective straight into the result, no padding and no concatenation
        out = local.new_empty((n_total,) + tuple(local.shape[1:]))
        dist.all_gather_into_tensor(out, local.contiguous(), group=group)
        return out
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


def allreduce_bucket(tensors: Sequence[torch.Tensor], group=None, async_op: bool = False):
    """Sum `tensors` over ranks with one flat all-reduce (in place).  With `async_op` the collective is only
    enqueued; the returned callable waits for it (on the current stream) and copies the sums back."""
    if not tensors or not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1:
        return (lambda: None) if async_op else None
    flat = flatten_bucket(tensors)
    work = dist.all_reduce(flat, op=dist.ReduceOp.SUM, group=group, async_op=async_op)

    def finish():
        if work is not None:
            work.wait()
        for t, v in zip(tensors, unflatten_bucket(flat, tensors)):
            t.copy_(v)
    if async_op:
        return finish
    finish()
    return None


def broadcast_parameters(network: torch.nn.Module, src: int = 0, group=None) -> None:
    """Overwrite every rank's parameters with rank `src`'s.  Replicas that apply the same summed gradients stay
    identical only as long as every replicated kernel is deterministic; call this at start-up (after loading a
    checkpoint) and every few thousand steps to cancel any drift."""
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1:
        return
    with torch.no_grad():
        for p in network.parameters():
            if p.is_contiguous():
                dist.broadcast(p, src=src, group=group)
            else:                                       # channels_last_3d decoder weights: broadcast the storage order
                flat = torch.as_strided(p, (p.numel(),), (1,)) if p.dim() == 5 and \
                    p.is_contiguous(memory_format=torch.channels_last_3d) else p.contiguous().view(-1)
                dist.broadcast(flat, src=src, group=group)
                if flat.data_ptr() != p.data_ptr():
                    p.copy_(flat.view_as(p))


def max_parameter_divergence(network: torch.nn.Module, group=None) -> float:
    """max over parameters and ranks of |p_rank - p_rank0| (0.0 = the replicas are bit-identical)."""
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1:
        return 0.0
    worst = torch.zeros(1, device=next(network.parameters()).device)
    with torch.no_grad():
        for p in network.parameters():
            ref = p.detach().clone()
            dist.broadcast(ref, src=0, group=group)
            worst = torch.maximum(worst, (p.detach() - ref).abs().max().reshape(1))
    dist.all_reduce(worst, op=dist.ReduceOp.MAX, group=group)
    return float(worst)


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
    run before every eager exchange with HN_CHECK_SAME_FRAME=1) - with per-rank frames use an all-reduce of the
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
        """Raise unless the boundary tensors (functions of the frame and the replicated parameters only) agree on
        every rank: max and min over ranks of a per-tensor checksum within 1e-5 relative (with `cfg.b200.decoder = 'cudnn'` the decoder is not
        bit-reproducible between processes; a different frame moves the checksums by orders of magnitude more)."""
        if not self.boundary or not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1:
            return
        sums = torch.stack([t.detach().double().abs().sum() for t in self.boundary.values()])
        pair = torch.stack([sums, -sums])
        dist.all_reduce(pair, op=dist.ReduceOp.MAX, group=group)
        if bool(((pair[0] + pair[1]) > 1e-5 * pair[0].abs() + 1e-12).any()):
            raise RuntimeError('CutBReducer: ranks hold different frames / front-end outputs in this step; the cut-B '
                               'gradient exchange needs the same frame on every rank (use cut A otherwise)')

    def leaf_inputs(self) -> List[torch.Tensor]:
        return self.mlp_params + list(self.boundary.values())

    def exchange_and_finish(self, loss: torch.Tensor, group=None, overlap: bool = True) -> int:
        """Backward to the boundary, the two bucketed all-reduces, then the replicated tail.
        Returns the number of floats exchanged."""
        if os.environ.get('HN_CHECK_SAME_FRAME', '0') == '1' and not torch.cuda.is_current_stream_capturing():
            self.check_same_frame(group)
        from humannerf_b200 import ops
        leaves = self.leaf_inputs()
        n_mlp = len(self.mlp_params)
        b_tensors = list(self.boundary.values())
        # The field's backward runs only its DATA phase here (ops.deferred_weight_gradients): the boundary gradients
        # are complete after the motion field's backward, their all-reduce is enqueued on the communicator's stream,
        # and only then the 1.7 ms weight-gradient kernel is launched - the collective on the critical path hides
        # under it.  The MLP bucket (which that kernel fills) follows on the communicator's stream and is joined
        # before the parameters' .grad is set, so it overlaps the decoder / FK / pose-decoder backward.
        with ops.deferred_weight_gradients():
            grads = torch.autograd.grad(loss, leaves, retain_graph=True, allow_unused=True)
            grads = [torch.zeros_like(t) if g is None else g.contiguous() for g, t in zip(grads, leaves)]
            b_grads = grads[n_mlp:]
            finish_boundary = allreduce_bucket(b_grads, group=group, async_op=overlap)
            ops.flush_weight_gradients()
        if finish_boundary is not None:
            finish_boundary()
        finish_mlp = allreduce_bucket(grads[:n_mlp], group=group, async_op=overlap)
        live = [(t, g) for t, g in zip(b_tensors, b_grads) if t.requires_grad]
        if live:
            torch.autograd.backward([t for t, _ in live], [g for _, g in live])
        if finish_mlp is not None:
            finish_mlp()
        for p, g in zip(self.mlp_params, grads[:n_mlp]):
            if p.grad is None:
                p.grad = g
            else:
                p.grad.add_(g)          # in place: keeps the flat-buffer views of training.FusedAdam
        return int(sum(g.numel() for g in grads))
