"""Worker of tests/test_gpu_multi.py (launched with torch.distributed.run, world size 2 or more).

Checks on the REAL Network (SURVEY 8e, cfg3 split):
  1. one fp32 step of tests/golden/train_step_cfg2.npz (p20: 2400 rays) split over the ranks - local slice, rgb
     all_gather, identical loss, cut-B exchange (parallel.CutBReducer) - gives the gradients of the same step run in a
     single process (all 55 tensors);
  2. the loss every rank evaluates is the single-process loss;
  3. after 3 bf16 steps (one CUDA graph each with NCCL; eager with gloo) the replicas' parameters are still identical.
Backend: NCCL with one GPU per rank when the box has enough GPUs, otherwise gloo with every rank on cuda:0 (the
transport changes, the arithmetic under test does not).  Prints one line `MULTI {json}` on rank 0."""
import json
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
sys.path.insert(0, os.path.dirname(HERE))
import step_fixture as sf  # noqa: E402

from humannerf_b200 import synthetic as syn  # noqa: E402


def main():
    from humannerf_b200.config import make_cfg
    from humannerf_b200.network import Network
    from humannerf_b200.parallel import (CutBReducer, GatherRowsWithGrad, broadcast_parameters,
                                         max_parameter_divergence, shard_bounds)
    from humannerf_b200.training import GraphedTrainStep, PerceptualLoss, RenderLoss, make_optimizer
    world, rank, local = int(os.environ['WORLD_SIZE']), int(os.environ['RANK']), int(os.environ['LOCAL_RANK'])
    nccl = torch.cuda.device_count() >= world
    dev = torch.device('cuda', local if nccl else 0)
    torch.cuda.set_device(dev)
    dist.init_process_group('nccl' if nccl else 'gloo', **({'device_id': dev} if nccl else {}))
    T = lambda x: torch.from_numpy(np.ascontiguousarray(x)).to(dev)

    st = sf.load_step('p20')
    cfg = make_cfg()
    cfg.perturb = 1.0
    cfg.b200.heavy_outputs = False
    cfg.b200.mlp_precision = 'fp32'
    net = Network()
    net.load_state_dict(syn.fixture_state_dict(syn.REFERENCE_PARAM_SHAPES, dense=True), strict=True)
    net = net.to(dev).train()
    crit = RenderLoss(sf.make_perceptual(st)).to(dev)
    fr = {k: T(v) for k, v in st['frame'].items()}
    kw = {k: fr[k] for k in ('dst_Rs', 'dst_Ts', 'cnl_gtfms', 'motion_weights_priors', 'dst_posevec',
                             'cnl_bbox_min_xyz', 'cnl_bbox_max_xyz', 'cnl_bbox_scale_xyz')}
    kw['bgcolor'] = T(st['bgcolor'])
    bg01 = kw['bgcolor'] / 255.0
    rays, near, far, t_rand = T(st['rays']), T(st['near']), T(st['far']), st['t_rand'].to(dev)
    index, targets = T(st['scatter_index']), T(st['targets'])
    R = near.shape[0]
    it = torch.full((1,), st['iter'], device=dev)

    with torch.backends.cudnn.flags(enabled=True, allow_tf32=False):
        # ---- single process: the whole step on this rank
        net.zero_grad()
        out = net(rays=rays, near=near, far=far, iter_val=it, t_rand=t_rand, **kw)
        loss1 = crit(out['rgb'], None, bg01, targets, index=index)['total']
        loss1.backward()
        ref = {k: p.grad.detach().clone() for k, p in net.named_parameters()}
        # ---- the same step split over the ranks
        net.zero_grad(set_to_none=True)
        a, b = shard_bounds(R, rank, world)
        out = net(rays=rays[:, a:b], near=near[a:b], far=far[a:b], iter_val=it, t_rand=t_rand[a:b], **kw)
        rgb = GatherRowsWithGrad.apply(out['rgb'], R)
        lossN = crit(rgb, None, bg01, targets, index=index)['total']
        red = CutBReducer(net)
        red.set_boundary(**net.boundary_tensors)
        red.check_same_frame()
        n_floats = red.exchange_and_finish(lossN)
    rows = {}
    for k, p in net.named_parameters():
        rows[k] = float((p.grad.double() - ref[k].double()).norm() / (ref[k].double().norm() + 1e-30))
    worst = {}
    for k, v in rows.items():
        worst[k.split('.')[0]] = max(worst.get(k.split('.')[0], 0.0), v)

    loss_single, loss_split = float(loss1), float(lossN)
    del out, rgb, lossN, loss1          # the eager graphs above must not outlive this point (see GraphedTrainStep._capture)
    # ---- replicas stay identical through optimizer steps (bf16 tensor-core path, strong split)
    cfg.b200.mlp_precision = 'bf16'
    cfg.b200.jitter_rng = 'device'
    cfg.perturb = 0.0
    broadcast_parameters(net)
    opt = make_optimizer(net)
    crit2 = RenderLoss(PerceptualLoss(0)).to(dev)
    cap = 6 * st['patch'] ** 2
    pad = cap - R
    batch = {'rays': torch.cat([rays, rays.new_zeros(3, pad, 3)], 1), 'near': torch.cat([near, near.new_zeros(pad, 1)]),
             'far': torch.cat([far, far.new_zeros(pad, 1)]), 'scatter_index': torch.cat([index, index.new_full((pad,), -1)]),
             'target_patches': targets}
    losses = []
    if nccl:
        step = GraphedTrainStep(net, crit2, opt, kw, bg01, 6, st['patch'], reducer=red, world=world, scaling='strong',
                                rank=rank)
        for s in range(3):
            losses.append(float(step(batch, st['iter'] + s)))
        mode = f'cuda graph ({step.launches_per_step} launches of this library per step)'
        step.graph = None
    else:
        for s in range(3):
            opt.zero_grad(set_to_none=True)
            a, b = shard_bounds(cap, rank, world)
            o = net(rays=batch['rays'][:, a:b], near=batch['near'][a:b], far=batch['far'][a:b], iter_val=st['iter'] + s, **kw)
            full = GatherRowsWithGrad.apply(o['rgb'], cap)
            ls = crit2(full, None, bg01, targets, index=batch['scatter_index'])['total']
            red.set_boundary(**net.boundary_tensors)
            red.exchange_and_finish(ls)
            opt.step()
            losses.append(float(ls))
        mode = 'eager (gloo cannot be captured)'
    div = max_parameter_divergence(net)
    sd0 = syn.fixture_state_dict(syn.REFERENCE_PARAM_SHAPES, dense=True)
    moved = float(max((p.detach() - sd0[k].to(dev)).abs().max() for k, p in net.state_dict().items()))
    torch.cuda.synchronize()
    dist.barrier()
    if rank == 0:
        print('MULTI ' + json.dumps({
            'world': world, 'backend': 'nccl' if nccl else 'gloo', 'devices': world if nccl else 1,
            'loss_single': loss_single, 'loss_split': loss_split, 'floats_exchanged': n_floats,
            'grad_rel_l2_worst_by_group': worst, 'grad_rel_l2_worst': max(rows.values()),
            'steps_mode': mode, 'step_losses': losses, 'param_divergence_after_3_steps': div,
            'params_moved_max_abs': moved}), flush=True)
    sys.stdout.flush()
    os._exit(0)          # skip communicator destructors (a captured NCCL graph hung a 2-GPU exit in round 1)


if __name__ == '__main__':
    main()
