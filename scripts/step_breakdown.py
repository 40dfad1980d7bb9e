"""Where does a train step go?  Runs the bench's train step under torch.profiler and prints (a) wall
time per step, (b) GPU-busy time per step, (c) the top kernels by total GPU time, (d) CPU-side time of
the step's phases.  Diagnostic only; numbers under a profiler are never bench values."""
import argparse
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from humannerf_b200 import ops, synthetic as syn          # noqa: E402
from humannerf_b200.config import make_cfg               # noqa: E402
from humannerf_b200.data import make_train_batch         # noqa: E402
from humannerf_b200.network import Network               # noqa: E402
from humannerf_b200.training import PerceptualLoss, RenderLoss, make_optimizer   # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument('--steps', type=int, default=8)
ap.add_argument('--precision', default='bf16')
ap.add_argument('--no-profiler', action='store_true')
args = ap.parse_args()
H = W = 512
dev = torch.device('cuda:0')
cfg = make_cfg()
cfg.perturb = 1.0
cfg.b200.jitter_rng = 'device'
cfg.b200.heavy_outputs = False
cfg.b200.mlp_precision = args.precision
net = Network()
net.load_state_dict(syn.fixture_state_dict(syn.REFERENCE_PARAM_SHAPES, dense=True))
net = net.to(dev).train()
opt = make_optimizer(net)
crit = RenderLoss(PerceptualLoss(0)).to(dev)
frame = syn.make_frame(0)
fr = {k: torch.from_numpy(np.ascontiguousarray(v)).to(dev) for k, v in frame.items()}
K, E = syn.look_at_camera(W, H, 1250.0)
o64, d64 = ops.gen_rays(H, W, K, E, dev)
near64, far64, mask = ops.rays_bbox(np.stack([frame['dst_bbox_min_xyz'], frame['dst_bbox_max_xyz']]), o64, d64)
o_c, d_c, near_c, far_c, pix = ops.compact_rays(mask, o64, d64, near64, far64)
kw = {k: fr[k] for k in ('dst_Rs', 'dst_Ts', 'cnl_gtfms', 'motion_weights_priors', 'dst_posevec', 'cnl_bbox_min_xyz',
                         'cnl_bbox_max_xyz', 'cnl_bbox_scale_xyz', 'bgcolor')}
mask_np = mask.cpu().numpy().astype(bool)
yy, xx = np.mgrid[0:H, 0:W]
subject = ((((xx - W / 2) / (0.16 * W)) ** 2 + ((yy - H / 2) / (0.38 * H)) ** 2) < 1.0) & mask_np.reshape(H, W)
image = np.random.RandomState(1).rand(H, W, 3).astype(np.float32)
rng = np.random.RandomState(0)
o_np, d_np = o_c.double().cpu().numpy(), d_c.double().cpu().numpy()
near_np, far_np = near_c[:, 0].cpu().numpy(), far_c[:, 0].cpu().numpy()
batches = []
for _ in range(args.steps + 3):
    b = make_train_batch(o_np, d_np, near_np, far_np, mask_np, subject, image, H, W, 6, 32, rng)
    batches.append({k: torch.from_numpy(np.ascontiguousarray(v)).to(dev) for k, v in b.items() if k != 'patch_div_indices'})
bg01 = fr['bgcolor'] / 255.0
phase = {}


def tick(name, t0):
    phase[name] = phase.get(name, 0.0) + time.perf_counter() - t0
    return time.perf_counter()


def step(b):
    t = time.perf_counter()
    opt.zero_grad(set_to_none=True)
    t = tick('zero_grad', t)
    out = net(rays=b['rays'], near=b['near'], far=b['far'], iter_val=60000, **kw)
    t = tick('forward', t)
    losses = crit(out['rgb'], b['patch_masks'], bg01, b['target_patches'])
    t = tick('loss', t)
    losses['total'].backward()
    t = tick('backward', t)
    opt.step()
    t = tick('adam', t)


for i in range(3):
    step(batches[i])
torch.cuda.synchronize()
phase.clear()
t0 = time.perf_counter()
if args.no_profiler:
    for i in range(args.steps):
        step(batches[3 + i])
    torch.cuda.synchronize()
    wall = time.perf_counter() - t0
    print(f'wall per step {1e3 * wall / args.steps:.3f} ms (no profiler)')
    print('CPU-side phase time per step (ms):', {k: round(1e3 * v / args.steps, 3) for k, v in phase.items()})
    sys.exit(0)
from torch.profiler import ProfilerActivity, profile  # noqa: E402
with profile(activities=[ProfilerActivity.CPU, ProfilerActivity.CUDA]) as prof:
    for i in range(args.steps):
        step(batches[3 + i])
    torch.cuda.synchronize()
wall = time.perf_counter() - t0
print(f'wall per step {1e3 * wall / args.steps:.3f} ms (under profiler)')
print('CPU-side phase time per step (ms):', {k: round(1e3 * v / args.steps, 3) for k, v in phase.items()})
kern = {}
for e in prof.events():
    if e.device_type == torch.autograd.DeviceType.CUDA:
        k = kern.setdefault(e.name, [0.0, 0])
        k[0] += e.device_time if hasattr(e, 'device_time') else e.cuda_time
        k[1] += 1
tot = sum(v[0] for v in kern.values())
print(f'GPU busy per step {tot / args.steps / 1e3:.3f} ms, {sum(v[1] for v in kern.values()) / args.steps:.0f} launches per step')
for name, (us, n) in sorted(kern.items(), key=lambda kv: -kv[1][0])[:45]:
    print(f'{us / args.steps / 1e3:9.3f} ms/step {n / args.steps:7.1f} calls/step  {name[:110]}')
