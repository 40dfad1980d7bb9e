"""Short command for ncu: renders `--rays` rays x 128 samples on the bf16 tcgen05 path (and,
with --train, one train-style forward / backward / Adam step) so a capture replays few, representative launches."""
import argparse
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from humannerf_b200 import ops, synthetic as syn          # noqa: E402
from humannerf_b200.config import make_cfg               # noqa: E402
from humannerf_b200.network import Network               # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument('--rays', type=int, default=32768)
ap.add_argument('--train', action='store_true')
ap.add_argument('--reps', type=int, default=2)
ap.add_argument('--precision', default='bf16', choices=['bf16', 'fp32'], help='MLP arithmetic of the --train pass')
args = ap.parse_args()
dev = 'cuda:0'
cfg = make_cfg()
cfg.b200.heavy_outputs = False
cfg.b200.jitter_rng = 'device'
net = Network()
net.load_state_dict(syn.fixture_state_dict(syn.REFERENCE_PARAM_SHAPES, dense=True))
net = net.to(dev)
from humannerf_b200.training import make_optimizer      # noqa: E402
opt = make_optimizer(net) if args.train else None        # FusedAdam: one adam_kernel launch per step
frame = syn.make_frame(0)
fr = {k: torch.from_numpy(np.ascontiguousarray(v)).to(dev) for k, v in frame.items()}
K, E = syn.look_at_camera(512, 512, 1250.0)
o64, d64 = ops.gen_rays(512, 512, K, E, dev)
near, far, mask = ops.rays_bbox(np.stack([frame['dst_bbox_min_xyz'], frame['dst_bbox_max_xyz']]), o64, d64)
o, d, nr, fr_, _ = ops.compact_rays(mask, o64, d64, near, far)
R = min(args.rays, o.shape[0])
kw = {k: fr[k] for k in ('dst_Rs', 'dst_Ts', 'cnl_gtfms', 'motion_weights_priors', 'dst_posevec', 'cnl_bbox_min_xyz',
                         'cnl_bbox_max_xyz', 'cnl_bbox_scale_xyz', 'bgcolor')}
rays = torch.stack([o[:R], d[:R], d[:R]], 0)
for _ in range(args.reps):
    if args.train:
        cfg.b200.mlp_precision, cfg.perturb = args.precision, 1.0
        net.train()
        opt.zero_grad()
        out = net(rays=rays, near=nr[:R], far=fr_[:R], iter_val=60000, **kw)
        (out['rgb'] ** 2).mean().backward()
        opt.step()
    else:
        cfg.b200.mlp_precision, cfg.perturb = 'bf16', 0.0
        net.eval()
        with torch.no_grad():
            out = net(rays=rays, near=nr[:R], far=fr_[:R], iter_val=cfg.eval_iter, **kw)
torch.cuda.synchronize()
print('ok', float(out['rgb'].sum()))
