"""Forward-render measurements at the shapes of BASELINE.json configs 1, 4 and 5:

  cfg1  512 x 512, 128 samples/ray, one pose
  cfg4  freeview: 1024 x 1024 views of one pose on a 360-view orbit (camera_util.py:84-109); a few of the
        360 views are rendered and the per-view time is reported (views are independent: time scales linearly)
  cfg5  movement: 1920 x 1080 frames of a smoothed random pose sequence, 256 samples/ray

For every frame the whole product path runs on the device: rays from (K, E) (hn_gen_rays), bbox
intersection + compaction (hn_rays_bbox / hn_compact_rays), Network.forward on the bf16 tensor-core path,
scatter into the H x W image.  With WORLD_SIZE > 1 (torchrun) the rays of a frame are sharded over the
ranks and gathered (20 B/ray).  Timed with CUDA events, max over ranks; dense and with opt-in empty-space
skipping (eps = 1e-7, |error| <= 2 * S * eps).  One JSON line per configuration.
"""
import argparse
import json
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from humannerf_b200 import ops, synthetic as syn                      # noqa: E402
from humannerf_b200.config import make_cfg                            # noqa: E402
from humannerf_b200.network import Network                            # noqa: E402
from humannerf_b200.parallel import all_gather_rows, shard_bounds     # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument('--configs', default='1,4,5')
ap.add_argument('--frames', type=int, default=3, help='frames / views rendered per configuration')
args = ap.parse_args()

world = int(os.environ.get('WORLD_SIZE', '1'))
rank = int(os.environ.get('RANK', '0'))
local = int(os.environ.get('LOCAL_RANK', '0'))
torch.cuda.set_device(local)
dev = torch.device('cuda', local)
if world > 1:
    dist.init_process_group('nccl', device_id=dev)

cfg = make_cfg()
cfg.b200.heavy_outputs = False
cfg.b200.mlp_precision = 'bf16'
cfg.b200.cache_eval_volume = True
cfg.perturb = 0.0
cfg.chunk = 1 << 30                           # one pass per frame: 180 GB of HBM hold every per-sample tensor
net = Network()
net.load_state_dict(syn.fixture_state_dict(syn.REFERENCE_PARAM_SHAPES, dense=True))
net = net.to(dev).eval()

CONFIGS = {
    1: dict(name='cfg1 512x512 S=128', W=512, H=512, focal=1250.0, S=128, orbit=False, poses=False),
    4: dict(name='cfg4 freeview 1024x1024 S=128 (orbit views)', W=1024, H=1024, focal=2500.0, S=128, orbit=True, poses=False),
    5: dict(name='cfg5 movement 1920x1080 S=256 (pose sequence)', W=1920, H=1080, focal=2637.0, S=256, orbit=False, poses=True),
}


def frame_tensors(seed):
    f = syn.make_frame(seed)
    return f, {k: torch.from_numpy(np.ascontiguousarray(v)).to(dev) for k, v in f.items()}


def render(frame, fr, K, E, W, H, cull):
    """One frame through the product path; returns (image (H*W,3), n_valid rays)."""
    cfg.b200.cull_mask_eps = cull
    bounds = np.stack([frame['dst_bbox_min_xyz'], frame['dst_bbox_max_xyz']])
    o64, d64 = ops.gen_rays(H, W, K, E, dev)
    near, far, mask = ops.rays_bbox(bounds, o64, d64)
    o, d, nr, fa, pix = ops.compact_rays(mask, o64, d64, near, far)
    n = o.shape[0]
    a, b = shard_bounds(n, rank, world)
    kw = {k: fr[k] for k in ('dst_Rs', 'dst_Ts', 'cnl_gtfms', 'motion_weights_priors', 'dst_posevec', 'cnl_bbox_min_xyz',
                             'cnl_bbox_max_xyz', 'cnl_bbox_scale_xyz', 'bgcolor')}
    out = net(rays=torch.stack([o[a:b], d[a:b], d[a:b]], 0), near=nr[a:b], far=fa[a:b], iter_val=cfg.eval_iter, **kw)
    packed = torch.cat([out['rgb'], out['alpha'][:, None], out['depth'][:, None]], dim=1)
    full = all_gather_rows(packed, n) if world > 1 else packed
    image = ops.scatter_rows(full[:, :3], pix.long(), torch.zeros(H * W, 3, device=dev))
    return image, full, n


with torch.no_grad():
    for cid in [int(c) for c in args.configs.split(',')]:
        c = CONFIGS[cid]
        cfg.N_samples = c['S']
        K, E0 = syn.look_at_camera(c['W'], c['H'], c['focal'])
        frames = []
        for i in range(args.frames):
            E = syn.rotate_camera_y(E0, 2 * np.pi * (i * 360 // max(args.frames, 1)) / 360.0) if c['orbit'] else E0.astype(np.float64)
            frames.append((frame_tensors(i if c['poses'] else 0), E))
        res = {}
        for mode, cull in (('dense', 0.0), ('skip_1e-7', 1e-7)):
            for (frame, fr), E in frames[:1]:
                render(frame, fr, K.astype(np.float64), E, c['W'], c['H'], cull)      # warm-up
            torch.cuda.synchronize()
            if world > 1:
                dist.barrier()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            rays = 0
            for (frame, fr), E in frames:
                image, full, n = render(frame, fr, K.astype(np.float64), E, c['W'], c['H'], cull)
                rays += n
            e1.record()
            torch.cuda.synchronize()
            ms = e0.elapsed_time(e1)
            if world > 1:
                t = torch.tensor([ms], device=dev, dtype=torch.float64)
                dist.all_reduce(t, op=dist.ReduceOp.MAX)
                ms = float(t[0])
            assert torch.isfinite(full).all() and float(full[:, 3].min()) >= 0.0 and float(full[:, 3].max()) < 1.5
            res[mode] = {'ms_per_frame': ms / len(frames), 'rays_per_sec': rays / (ms * 1e-3),
                         'samples_per_sec': rays * c['S'] / (ms * 1e-3), 'valid_rays_per_frame': rays / len(frames),
                         'pixels_per_sec': c['W'] * c['H'] * len(frames) / (ms * 1e-3)}
            if mode != 'dense':
                res[mode]['samples_skipped_frac_last_frame_rank0'] = float(net.last_cull_fraction)
            res[mode + '_alpha_mean'] = float(full[:, 3].mean())
        if rank == 0:
            print(json.dumps({'config': c['name'], 'n_gpus': world, 'frames_rendered': len(frames), **res}), flush=True)

if world > 1:
    torch.cuda.synchronize()
    dist.barrier()
    sys.stdout.flush()
    os._exit(0)
