"""Times the forward render legs (CUDA events) for a given number of rays; prints per-op ms."""
import argparse, os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from humannerf_b200 import ops, synthetic as syn
from humannerf_b200.config import make_cfg
from humannerf_b200.network import Network
ap = argparse.ArgumentParser(); ap.add_argument('--rays', type=int, default=232224); ap.add_argument('--reps', type=int, default=5); ap.add_argument('--cull', type=float, default=0.0); ap.add_argument('--chunk', type=int, default=0); ap.add_argument('--cache-vol', action='store_true')
args = ap.parse_args()
dev = 'cuda:0'
cfg = make_cfg(); cfg.b200.heavy_outputs = False; cfg.b200.mlp_precision = 'bf16'; cfg.perturb = 0.0; cfg.b200.cull_mask_eps = args.cull
if args.chunk: cfg.chunk = args.chunk
cfg.b200.cache_eval_volume = args.cache_vol
net = Network(); net.load_state_dict(syn.fixture_state_dict(syn.REFERENCE_PARAM_SHAPES, dense=True)); net = net.to(dev).eval()
frame = syn.make_frame(0)
fr = {k: torch.from_numpy(np.ascontiguousarray(v)).to(dev) for k, v in frame.items()}
K, E = syn.look_at_camera(512, 512, 1250.0)
o64, d64 = ops.gen_rays(512, 512, K, E, dev)
near, far, mask = ops.rays_bbox(np.stack([frame['dst_bbox_min_xyz'], frame['dst_bbox_max_xyz']]), o64, d64)
o, d, nr, fr_, _ = ops.compact_rays(mask, o64, d64, near, far)
R = min(args.rays, o.shape[0])
kw = {k: fr[k] for k in ('dst_Rs', 'dst_Ts', 'cnl_gtfms', 'motion_weights_priors', 'dst_posevec', 'cnl_bbox_min_xyz', 'cnl_bbox_max_xyz', 'cnl_bbox_scale_xyz', 'bgcolor')}
rays = torch.stack([o[:R], d[:R], d[:R]], 0)
with torch.no_grad():
    cfg.b200.cull_mask_eps = 0.0
    ref = net(rays=rays, near=nr[:R], far=fr_[:R], iter_val=cfg.eval_iter, **kw)
    cfg.b200.cull_mask_eps = args.cull
    for _ in range(2):
        net(rays=rays, near=nr[:R], far=fr_[:R], iter_val=cfg.eval_iter, **kw)
    torch.cuda.synchronize()
    ops.timing_start()
    for _ in range(args.reps):
        net(rays=rays, near=nr[:R], far=fr_[:R], iter_val=cfg.eval_iter, **kw)
    t = ops.timing_stop()
P = R * 128
ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
with torch.no_grad():
    ev0.record()
    for _ in range(args.reps):
        out = net(rays=rays, near=nr[:R], far=fr_[:R], iter_val=cfg.eval_iter, **kw)
    ev1.record(); torch.cuda.synchronize()
print(f'frame {ev0.elapsed_time(ev1) / args.reps:.2f} ms, culled fraction {net.last_cull_fraction:.3f}, max |rgb - dense| '
      f'{float((out["rgb"] - ref["rgb"]).abs().max()):.2e}, max |alpha - dense| {float((out["alpha"] - ref["alpha"]).abs().max()):.2e}')
print({k: round(v[0] / args.reps, 3) for k, v in t.items()}, 'TFLOP/s', round(1184768.0 * P / (t['field_fwd_bf16'][0] / args.reps * 1e-3) / 1e12, 1))
