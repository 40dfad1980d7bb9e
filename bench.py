#!/usr/bin/env python
"""Benchmark of the HumanNeRF per-ray rendering hot path on B200.

  python bench.py --gpus N --steps K --warmup W            (N>1: launched with torchrun)
  python bench.py --impl reference ...                      (CPU arm: the oracle port of the reference)
  python bench.py --workload cfg4|cfg5 --gpus N ...         (render workloads of BASELINE.json configs[3] / [4])

Default workload (BASELINE.json configs[1]/[2]): one training step = 6 patches of 32x32 rays of a
synthetic 512x512 frame (<= 6144 rays x 128 samples): patch selection -> forward -> loss (0.2 MSE + 1.0 LPIPS-VGG16, the
reference's LPIPS(net='vgg', pnet_rand=True) under seed 0) -> backward -> Adam, ZJU-387 adventure.yaml network shapes,
fixture ("dense") weights, synthetic 24-joint pose; MLPs on the bf16 tcgen05 kernels (--precision fp32 selects the SIMT
parity path).  The whole step - including the device patch selection (a3) - is one CUDA-graph replay
(training.GraphedTrainStep); per step the host supplies the reference's two randoms per patch and the schedule scalars.

`value` = rays/s of the dense step with inputs resident in HBM.  With N GPUs `value` is WEAK scaling (every rank steps on
its own 6 patches of the same frame, gradients averaged); the JSON line also carries `strong` = SURVEY 8e's split of ONE
6144-ray step over the ranks (rgb all_gather, identical loss, gradients summed).  Both use the cut-B gradient exchange
(parallel.CutBReducer) captured inside the graph.  `e2e` = the same step from the host: numpy draws of the patch randoms
(the reference's RNG calls), pinned staging, H2D, replay, loss read-back.  Further keys: the roofline of the MLP kernel
group (per-op CUDA events of an eager re-run of the same steps; DRAM traffic from the committed ncu capture named in
profiles/traffic.json), the step with opt-in empty-space skipping, the forward-render leg, the CPU baseline (oracle port,
whole step incl. LPIPS and Adam, bounded sample) and the eager-PyTorch-on-this-GPU baseline (the same oracle on cuda:0).
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

FLOPS_FWD = 1184768.0          # MLP forward FLOPs / sample (unpadded, 2 * MAC)
FLOPS_FWD_BWD = 3554304.0
MF_BYTES = 124.0               # motion field algorithmic bytes / sample (fp32, with 24-ch weights)
MF_BYTES_NOW = 28.0            # ... when the weights output is omitted
MF_ONCE = 3146880.0
H = W = 512
N_PATCHES, PATCH, N_SAMPLES = 6, 32, 128


def load_peaks():
    p = os.path.join(ROOT, 'MEASURED_PEAKS.json')
    if os.path.isfile(p):
        with open(p) as f:
            d = json.load(f)
        return {'hbm_gbs': d['hbm_gbs'], 'tf': d['bf16_tflops'], 'tf_sustained': d.get('bf16_tflops_sustained', d['bf16_tflops']),
                'source': 'measured'}
    return {'hbm_gbs': 6650.0, 'tf': 1590.0, 'tf_sustained': 1400.0, 'source': 'fallback'}


def load_traffic():
    """DRAM bytes per launch of the dominant kernels, read from the committed ncu capture (profiles/traffic.json names
    the capture each number comes from); None when the file is absent."""
    p = os.path.join(ROOT, 'profiles', 'traffic.json')
    if os.path.isfile(p):
        with open(p) as f:
            return json.load(f)
    return None


class ClockSampler:
    """SM clock / throttle reasons sampled DURING the timed region: NVML from a background thread every
    10 ms (pynvml), falling back to polling `nvidia-smi` when NVML is not importable."""
    Q = ('clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,'
         'clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap')

    def __init__(self, index=0):
        self.index, self.sm, self.reasons, self.max_mhz = index, [], set(), None
        self.stop_flag, self.thread, self.source = False, None, 'nvml'
        try:
            import pynvml
            pynvml.nvmlInit()
            vis = os.environ.get('CUDA_VISIBLE_DEVICES')
            phys = int(vis.split(',')[index]) if vis and all(t.strip().isdigit() for t in vis.split(',')) else index
            self.nv, self.h = pynvml, pynvml.nvmlDeviceGetHandleByIndex(phys)
            self.max_mhz = float(pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM))
        except Exception:
            self.nv, self.source = None, 'nvidia-smi'

    def _run_nvml(self):
        nv = self.nv
        names = {'hw_slowdown': getattr(nv, 'nvmlClocksEventReasonHwSlowdown', 0x8),
                 'hw_thermal_slowdown': getattr(nv, 'nvmlClocksEventReasonHwThermalSlowdown', 0x40),
                 'sw_thermal_slowdown': getattr(nv, 'nvmlClocksEventReasonSwThermalSlowdown', 0x20),
                 'sw_power_cap': getattr(nv, 'nvmlClocksEventReasonSwPowerCap', 0x4)}
        get_reasons = getattr(nv, 'nvmlDeviceGetCurrentClocksEventReasons', None) or \
            getattr(nv, 'nvmlDeviceGetCurrentClocksThrottleReasons')
        while not self.stop_flag:
            try:
                self.sm.append(float(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM)))
                mask = int(get_reasons(self.h))
                self.reasons.update(n for n, bit in names.items() if mask & bit)
            except Exception:
                pass
            time.sleep(0.01)

    def _run_smi(self):
        names = ['hw_slowdown', 'hw_thermal_slowdown', 'sw_thermal_slowdown', 'sw_power_cap']
        while not self.stop_flag:
            try:
                out = subprocess.run(['nvidia-smi', f'--id={self.index}', f'--query-gpu={self.Q}',
                                      '--format=csv,noheader,nounits'], capture_output=True, text=True, timeout=5).stdout
                r = [c.strip() for c in out.strip().split('\n')[0].split(',')] if out.strip() else []
                if r and r[0].replace('.', '').isdigit():
                    self.sm.append(float(r[0]))
                if len(r) > 1 and r[1].replace('.', '').isdigit():
                    self.max_mhz = max(self.max_mhz or 0.0, float(r[1]))
                self.reasons.update(n for n, v in zip(names, r[2:6]) if v.lower().startswith('active'))
            except Exception:
                pass
            time.sleep(0.1)

    def start(self):
        self.thread = threading.Thread(target=self._run_nvml if self.nv else self._run_smi, daemon=True)
        self.thread.start()

    def stop(self):
        self.stop_flag = True
        if self.thread:
            self.thread.join(timeout=6)
        return {'sm_mhz': float(np.median(self.sm)) if self.sm else None, 'sm_min_mhz': min(self.sm) if self.sm else None,
                'sm_max_mhz': self.max_mhz, 'reasons': sorted(self.reasons), 'samples': len(self.sm), 'source': self.source}


# ----------------------------------------------------------------------------------------------
def scene_on_host(syn, seed=0):
    frame = syn.make_frame(seed)
    K, E = syn.look_at_camera(W, H, 1250.0)
    return frame, K, E


def workload_config(n_gpus, patch=None, note=''):
    patch = PATCH if patch is None else patch
    return {'workload': f'ZJU-387 adventure.yaml train step: 6 patches x {patch}x{patch} rays x 128 samples per GPU per step '
                        '(patch selection + fwd + 0.2 MSE + 1.0 LPIPS(VGG16, pnet_rand seed 0) + bwd + Adam), synthetic 512x512 '
                        'frame, fixture weights' + note,
            'rays_per_step_per_gpu': N_PATCHES * patch * patch, 'samples_per_ray': N_SAMPLES,
            'parallelism': f'dp{n_gpus} (value: weak - own patches per rank; strong: one step split by rays; 2 cut-B gradient '
                           'buckets per step)',
            'l2_policy': 'inputs change every step (new patches); activations '
                         f'{N_PATCHES * patch * patch * N_SAMPLES * 12.1e3 / 1e9:.1f} GB/step exceed the 126 MB L2'}


def oracle_train_steps(device, steps, warmup, patch, threads=None, rng_seed=0):
    """The oracle port of the reference's whole train step (patch selection, Network.forward, the reference's loss
    with LPIPS, backward, Adam) on `device`; returns (rays processed, seconds) over the timed steps.
    Used by the CPU arms (`--impl reference`, `cpu_baseline`) and by `eager_gpu_baseline` (device = cuda:0)."""
    import torch
    from humannerf_b200 import synthetic as syn
    from humannerf_b200.data import make_train_batch
    from humannerf_b200.training import PerceptualLoss
    from oracle import humannerf_oracle as orc
    if threads:
        torch.set_num_threads(threads)
    dev = torch.device(device)
    frame, K, E = scene_on_host(syn)
    ro, rd = syn.camera_rays_np(H, W, K, E)
    bounds = np.stack([frame['dst_bbox_min_xyz'], frame['dst_bbox_max_xyz']])
    near, far, mask, rd = orc.rays_bbox_near_far(bounds, ro, rd)
    ro_c, rd_c = ro[mask], rd[mask]
    yy, xx = np.mgrid[0:H, 0:W]
    subject = ((((xx - W / 2) / (0.16 * W)) ** 2 + ((yy - H / 2) / (0.38 * H)) ** 2) < 1.0) & mask.reshape(H, W)
    image = np.random.RandomState(1).rand(H, W, 3).astype(np.float32)
    t = lambda x: torch.from_numpy(np.ascontiguousarray(x)).to(dev)
    sd = {k: v.clone().to(dev).requires_grad_(True) for k, v in syn.fixture_state_dict(syn.REFERENCE_PARAM_SHAPES).items()}
    fr = {k: t(v) for k, v in frame.items()}
    lp = PerceptualLoss(0).to(dev)
    if dev.type == 'cuda':
        lp.forward = lambda x, y, _lp=lp: _torch_lpips(_lp, x, y)       # plain torch ops: no kernel of this repo
    opt = torch.optim.Adam(list(sd.values()), lr=5e-4)
    rng = np.random.RandomState(rng_seed)
    sync = (lambda: torch.cuda.synchronize(dev)) if dev.type == 'cuda' else (lambda: None)
    total, rays = 0.0, 0
    for step in range(warmup + steps):
        sync()
        t0 = time.perf_counter()
        b = make_train_batch(ro_c, rd_c, near, far, mask, subject, image, H, W, N_PATCHES, patch, rng)
        R = b['near'].shape[0]
        out = orc.network_forward(sd, t(b['rays']), t(b['near']), t(b['far']), fr, torch.full((1,), 60000, device=dev),
                                  n_samples=N_SAMPLES, t_rand=torch.rand(R, N_SAMPLES).to(dev))
        pm, tg = t(b['patch_masks']), t(b['target_patches'])
        img = orc.unpack_patches(out['rgb'], pm, fr['bgcolor'] / 255.0, b['patch_div_indices'])
        loss = 0.2 * torch.mean((img - tg) ** 2) + torch.mean(lp(img.permute(0, 3, 1, 2) * 2 - 1, tg.permute(0, 3, 1, 2) * 2 - 1))
        opt.zero_grad()
        loss.backward()
        opt.step()
        _ = float(loss.detach())
        sync()
        dt = time.perf_counter() - t0
        if step >= warmup:
            total += dt
            rays += R
    return rays, total


def _torch_lpips(lp, x, y):
    """training.PerceptualLoss evaluated with torch ops only (its CPU branch) on any device."""
    import torch
    n = x.shape[0]
    f = (torch.cat([x, y], 0) - lp.shift) / lp.scale
    total = 0.0
    for stage, lin in zip(lp.stages, lp.lins):
        f = stage(f)
        nf = f / (torch.sqrt(torch.sum(f ** 2, dim=1, keepdim=True) + 1e-10) + 1e-10)
        total = total + lin((nf[:n] - nf[n:]) ** 2).mean([2, 3], keepdim=True)
    return total


def reference_arm(args):
    """The reference's own CPU implementation of the path = the oracle port (the Python reference cannot travel to
    the GPU box), all host threads, the SAME train-step config as the GPU arm (6 patches of --ref-patch^2 rays,
    default 32), a bounded number of steps."""
    rank = int(os.environ.get('RANK', '0'))
    if rank != 0:
        return
    cores = os.cpu_count() or 1
    steps = max(1, min(args.steps, args.ref_max_steps))
    rays, total = oracle_train_steps('cpu', steps, args.warmup_ref, args.ref_patch, threads=cores)
    val = rays / total
    sample = (f'{steps} whole train steps (+{args.warmup_ref} warm-up) of {N_PATCHES} patches x {args.ref_patch}x{args.ref_patch} rays x '
              f'{N_SAMPLES} samples: patch selection + forward + MSE + LPIPS + backward + Adam, torch fp32 on {cores} threads')
    note = '' if args.ref_patch == PATCH else f' [CPU arm ran {args.ref_patch}x{args.ref_patch} patches]'
    print(json.dumps({
        'impl': 'reference', 'metric': 'train_step_rays_per_sec', 'value': val, 'unit': 'rays/s', 'n_gpus': args.gpus,
        'steps': steps, 'steps_requested': args.steps, 'warmup': args.warmup_ref, 'ms_per_step': 1e3 * total / steps,
        'higher_is_better': True, 'scaling': 'weak', 'vs_baseline': None, 'dtype': 'f32', 'data': 'synthetic',
        'config': workload_config(args.gpus, args.ref_patch, note),
        'cpu_baseline': {'value': val, 'unit': 'rays/s', 'cores': cores, 'kind': 'port', 'sample': sample},
        'e2e': {'value': val, 'unit': 'rays/s', 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0},
    }))


# ----------------------------------------------------------------------------------------------
RENDER_WORKLOADS = {
    'cfg4': dict(name='freeview render: 1024x1024 views of one synthetic pose on a 360-view orbit, 128 samples/ray '
                      '(BASELINE.json configs[3])', W=1024, H=1024, focal=2500.0, S=128, orbit=True, poses=False, total=360),
    'cfg5': dict(name='movement render: 1920x1080 frames of a synthetic pose sequence, 256 samples/ray '
                      '(BASELINE.json configs[4])', W=1920, H=1080, focal=2637.0, S=256, orbit=False, poses=True, total=600),
}


def render_workload(args):
    """cfg4 / cfg5: views (frames) are dealt round-robin to the ranks - zero communication (render).  A
    "step" is one view per rank through the whole device path (rays from K, E -> bbox -> compaction -> network ->
    image scatter); `value` = valid rays/s of the whole job; the per-view time extrapolates to the 360 / 600 frames."""
    import torch
    import torch.distributed as dist
    from humannerf_b200 import _lib, synthetic as syn
    from humannerf_b200.config import make_cfg
    from humannerf_b200.network import Network
    from humannerf_b200.render import render_view
    c = RENDER_WORKLOADS[args.workload]
    world, rank, local = int(os.environ.get('WORLD_SIZE', '1')), int(os.environ.get('RANK', '0')), int(os.environ.get('LOCAL_RANK', '0'))
    assert world == args.gpus, f'--gpus {args.gpus} but WORLD_SIZE={world} (launch N>1 with torchrun)'
    torch.cuda.set_device(local)
    dev = torch.device('cuda', local)
    if world > 1:
        dist.init_process_group('nccl', device_id=dev)
    cfg = make_cfg()
    cfg.b200.heavy_outputs, cfg.b200.mlp_precision, cfg.b200.cache_eval_volume = False, 'bf16', True
    cfg.perturb, cfg.N_samples = 0.0, c['S']
    cfg.chunk = 1 << 30                     # one pass per view: 180 GB of HBM hold every per-sample tensor
    net = Network()
    net.load_state_dict(syn.fixture_state_dict(syn.REFERENCE_PARAM_SHAPES, dense=True))
    net = net.to(dev).eval()
    K, E0 = syn.look_at_camera(c['W'], c['H'], c['focal'])
    K = K.astype(np.float64)

    def view(i):
        idx = i * world + rank                                            # round robin over the ranks
        frame = syn.make_frame(idx % 16 if c['poses'] else 0)
        fr = {k: torch.from_numpy(np.ascontiguousarray(v)).to(dev) for k, v in frame.items()}
        E = syn.rotate_camera_y(E0, 2 * np.pi * (idx % 360) / 360.0) if c['orbit'] else E0.astype(np.float64)
        return fr, E, np.stack([frame['dst_bbox_min_xyz'], frame['dst_bbox_max_xyz']])

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    views = [view(i) for i in range(args.warmup + args.steps)]          # host-side pose tensors, staged before timing
    res = {}
    with torch.no_grad():
        for mode, cull in (('dense', 0.0), ('skip_1e-7', 1e-7)):
            cfg.b200.cull_mask_eps = cull
            for fr, E, bounds in views[:args.warmup]:
                render_view(net, fr, K, E, c['H'], c['W'], bounds)
            barrier()
            clocks = ClockSampler(local)
            clocks.start()
            n0 = _lib.launch_count()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            rays = 0
            for fr, E, bounds in views[args.warmup:]:
                img, alpha, n = render_view(net, fr, K, E, c['H'], c['W'], bounds)
                rays += n
            e1.record()
            barrier()
            clk = clocks.stop()
            launches = _lib.launch_count() - n0
            stats = torch.tensor([e0.elapsed_time(e1), float(rays), float(launches)], device=dev, dtype=torch.float64)
            ms, rays_all, launches_all = float(stats[0]), float(stats[1]), float(stats[2])
            if world > 1:
                mx, sm = stats.clone(), stats.clone()
                dist.all_reduce(mx, op=dist.ReduceOp.MAX)
                dist.all_reduce(sm, op=dist.ReduceOp.SUM)
                ms, rays_all, launches_all = float(mx[0]), float(sm[1]), float(sm[2])
            assert torch.isfinite(img).all() and float(alpha.max()) < 1.5
            res[mode] = {'ms': ms, 'rays': rays_all, 'clocks': clk, 'launches': launches_all, 'skipped': float(net.last_cull_fraction)}
    d = res['dense']
    views_done = args.steps * world
    line = {
        'metric': 'render_rays_per_sec', 'value': d['rays'] / (d['ms'] * 1e-3), 'unit': 'rays/s', 'n_gpus': world, 'steps': args.steps,
        'warmup': args.warmup, 'ms_per_step': d['ms'] / args.steps, 'higher_is_better': True, 'scaling': 'weak',
        'vs_baseline': None, 'dtype': 'bf16', 'data': 'synthetic',
        'config': {'workload': c['name'] + '; views dealt round-robin to the ranks, no communication; step = one view per rank',
                   'views_per_step': world, 'resolution': [c['W'], c['H']], 'samples_per_ray': c['S'],
                   'l2_policy': f'every view is {d["rays"] / views_done * c["S"] * 52 / 1e9:.1f} GB of per-sample tensors (> 126 MB L2)'},
        'samples_per_sec': d['rays'] * c['S'] / (d['ms'] * 1e-3),
        'pixels_per_sec': c['W'] * c['H'] * views_done / (d['ms'] * 1e-3),
        'valid_rays_per_view': d['rays'] / views_done, 'ms_per_view_per_gpu': d['ms'] / args.steps,
        'whole_job_seconds_extrapolated': {'views': c['total'], 'dense': c['total'] / world * d['ms'] / args.steps / 1e3,
                                           'skip_1e-7': c['total'] / world * res['skip_1e-7']['ms'] / args.steps / 1e3},
        'empty_space_skipping': {'cull_mask_eps': 1e-7, 'rays_per_sec': res['skip_1e-7']['rays'] / (res['skip_1e-7']['ms'] * 1e-3),
                                 'ms_per_view_per_gpu': res['skip_1e-7']['ms'] / args.steps,
                                 'samples_skipped_frac_last_view': res['skip_1e-7']['skipped'], 'bound': 2 * c['S'] * 1e-7},
        'e2e': {'value': d['rays'] / (d['ms'] * 1e-3), 'unit': 'rays/s', 'h2d_bytes_per_step': 16 * 8 + 9 * 8 + 6 * 8,
                'd2h_bytes_per_step': 8, 'note': 'the timed region IS the public call: camera + bounds in from the host per view, '
                'ray count out (compaction size); pose tensors staged per view before timing'},
        'gpu_launches': int(d['launches']), 'clocks': d['clocks'],
        'roofline': None, 'cpu_baseline': None,
    }
    if rank == 0:
        print(json.dumps(line), flush=True)
    if world > 1:
        torch.cuda.synchronize()
        dist.barrier()
        sys.stdout.flush()
        os._exit(0)


# ----------------------------------------------------------------------------------------------
def main():
    global PATCH
    ap = argparse.ArgumentParser()
    ap.add_argument('--gpus', type=int, default=1)
    ap.add_argument('--steps', type=int, default=20)
    ap.add_argument('--warmup', type=int, default=5)
    ap.add_argument('--impl', default='ours', choices=['ours', 'reference'])
    ap.add_argument('--workload', default='train', choices=['train', 'cfg4', 'cfg5'],
                    help="train = BASELINE configs[1]/[2] (default); cfg4 / cfg5 = the render workloads of configs[3] / [4]")
    ap.add_argument('--ref-patch', type=int, default=PATCH, help='patch size of the CPU arm (default: the same 32x32 config)')
    ap.add_argument('--ref-max-steps', type=int, default=4, help='the CPU arm times at most this many steps (each ~10-15 s)')
    ap.add_argument('--warmup-ref', type=int, default=1)
    ap.add_argument('--no-cpu-baseline', action='store_true')
    ap.add_argument('--no-eager-gpu-baseline', action='store_true')
    ap.add_argument('--no-render', action='store_true')
    ap.add_argument('--no-skip-leg', action='store_true', help='skip the extra train leg with empty-space skipping')
    ap.add_argument('--no-strong-leg', action='store_true', help='N > 1: skip the strong-scaling leg')
    ap.add_argument('--patch', type=int, default=PATCH,
                    help='patch side of the train step: 32 = default.yaml / BASELINE.json cfg2 (default), 20 = what '
                         'zju_mocap/387/single_gpu.yaml sets (SURVEY 8d asks for both; >= 16 for the LPIPS trunk)')
    ap.add_argument('--eager', action='store_true', help='run the train step eagerly instead of as one CUDA graph')
    ap.add_argument('--precision', default='bf16', choices=['bf16', 'fp32'],
                    help='MLP arithmetic of the train step: bf16 tcgen05 kernels (default) or the fp32 SIMT parity path')
    args = ap.parse_args()
    assert args.patch >= 16, 'the LPIPS VGG16 trunk pools four times'
    PATCH = args.patch
    if args.impl == 'reference':
        return reference_arm(args)
    assert args.warmup >= 3, 'timing rules: at least 3 warm-up steps'
    if args.workload != 'train':
        return render_workload(args)

    import torch
    import torch.distributed as dist
    from humannerf_b200 import _lib, ops, synthetic as syn
    from humannerf_b200.config import make_cfg
    from humannerf_b200.data import DevicePatchSampler
    from humannerf_b200.network import Network
    from humannerf_b200.parallel import CutBReducer, all_gather_rows, broadcast_parameters, shard_bounds
    from humannerf_b200.render import frame_rays
    from humannerf_b200.training import GraphedTrainStep, PerceptualLoss, RenderLoss, make_optimizer

    if os.environ.get('NCCL_DEBUG', '').upper() == 'VERSION':
        os.environ['NCCL_DEBUG'] = 'WARN'          # keep stdout to the one JSON line
    world = int(os.environ.get('WORLD_SIZE', '1'))
    rank = int(os.environ.get('RANK', '0'))
    local = int(os.environ.get('LOCAL_RANK', '0'))
    assert world == args.gpus, f'--gpus {args.gpus} but WORLD_SIZE={world} (launch N>1 with torchrun)'
    torch.cuda.set_device(local)
    dev = torch.device('cuda', local)
    if world > 1:
        dist.init_process_group('nccl', device_id=dev)
    peaks = load_peaks()
    traffic = load_traffic()

    cfg = make_cfg()
    cfg.perturb = 1.0
    cfg.b200.jitter_rng = 'device'
    cfg.b200.heavy_outputs = False
    cfg.N_samples = N_SAMPLES
    if os.environ.get('HN_DECODER'):
        cfg.b200.decoder = os.environ['HN_DECODER']          # A/B switch: 'hn' or 'cudnn' (default: cfg.b200.decoder = 'auto')
    if os.environ.get('HN_FRONT_STREAMS'):
        cfg.b200.front_streams = os.environ['HN_FRONT_STREAMS'] != '0'      # A/B switch of the decoder side stream
    torch.manual_seed(0)
    net = Network()
    net.load_state_dict(syn.fixture_state_dict(syn.REFERENCE_PARAM_SHAPES, dense=True))
    net = net.to(dev)
    opt = make_optimizer(net)
    crit = RenderLoss(PerceptualLoss(0)).to(dev)

    # ---- scene: rays of the frame come from the product kernels (a1, a2, compaction)
    frame, K, E = scene_on_host(syn)
    fr = {k: torch.from_numpy(np.ascontiguousarray(v)).to(dev) for k, v in frame.items()}
    bounds = np.stack([frame['dst_bbox_min_xyz'], frame['dst_bbox_max_xyz']])
    o_c, d_c, near_c, far_c, pix, mask = frame_rays(H, W, K, E, bounds, dev)
    n_valid = o_c.shape[0]
    frame_kw = {k: fr[k] for k in ('dst_Rs', 'dst_Ts', 'cnl_gtfms', 'motion_weights_priors', 'dst_posevec',
                                   'cnl_bbox_min_xyz', 'cnl_bbox_max_xyz', 'cnl_bbox_scale_xyz', 'bgcolor')}

    # ---- forward render of the whole frame on the bf16 tensor-core path -> target image + subject mask
    def render_frame(sharded=False, cull_eps=0.0, one_chunk=False):
        """sharded: this rank renders its contiguous block of the frame's rays and one all_gather of
        (rgb, alpha, depth) = 20 B / ray assembles the frame on every rank (render)."""
        cfg.b200.mlp_precision, cfg.perturb = 'bf16', 0.0
        cfg.b200.cull_mask_eps = cull_eps
        cfg.b200.cache_eval_volume = True     # the decoded volume is pose independent: decoded once per weight version
        chunk0 = cfg.chunk
        if one_chunk:
            cfg.chunk = 1 << 30          # 180 GB of HBM: the whole frame's samples fit in one pass
        net.eval()
        a, b = shard_bounds(n_valid, rank, world) if (sharded and world > 1) else (0, n_valid)
        with torch.no_grad():
            rays = torch.stack([o_c[a:b], d_c[a:b], d_c[a:b]], 0)
            out = net(rays=rays, near=near_c[a:b], far=far_c[a:b], iter_val=cfg.eval_iter, **frame_kw)
            if sharded and world > 1:
                packed = torch.cat([out['rgb'], out['alpha'][:, None], out['depth'][:, None]], dim=1)
                full = all_gather_rows(packed, n_valid)
                out = {'rgb': full[:, :3], 'alpha': full[:, 3], 'depth': full[:, 4]}
        cfg.b200.mlp_precision, cfg.perturb = args.precision, 1.0
        cfg.b200.cull_mask_eps, cfg.chunk = 0.0, chunk0
        net.train()
        return out

    out = render_frame()
    image = torch.zeros(H * W, 3, device=dev)
    image[pix.long()] = out['rgb']
    alpha = torch.zeros(H * W, device=dev)
    alpha[pix.long()] = out['alpha']
    subject = (alpha > 0.5).reshape(H, W)
    if int(subject.sum()) < 100:
        yy, xx = torch.meshgrid(torch.arange(H, device=dev), torch.arange(W, device=dev), indexing='ij')
        subject = (((xx - W / 2) / (0.16 * W)) ** 2 + ((yy - H / 2) / (0.38 * H)) ** 2) < 1.0
    subject = subject & mask.reshape(H, W).bool()

    # ---- a3 on the device: the frame's rays / masks / image stay in HBM; per step the host draws the reference's randoms
    sampler = DevicePatchSampler(mask, subject, o_c, d_c, near_c, far_c, image.reshape(H, W, 3), H, W, N_PATCHES, PATCH)
    total_steps = args.warmup + args.steps
    bg01 = fr['bgcolor'] / 255.0
    reducer = CutBReducer(net)
    if world > 1:
        broadcast_parameters(net)

    def make_draws(seed, n):
        rng = np.random.RandomState(seed)
        return [sampler.draw(rng) for _ in range(n)]

    def eager_step(batch, it, scaling='weak'):
        opt.zero_grad(set_to_none=True)
        out = net(rays=batch['rays'], near=batch['near'], far=batch['far'], iter_val=it, **frame_kw)
        losses = crit(out['rgb'], None, bg01, batch['target_patches'], index=batch['scatter_index'])
        loss = losses['total'] / world
        if world > 1:
            reducer.set_boundary(**net.boundary_tensors)
            reducer.exchange_and_finish(loss)
        else:
            loss.backward()
        opt.step()
        return losses['total']

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(v):
        if world == 1:
            return float(v)
        t = torch.tensor([float(v)], device=dev, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t[0])

    it_val = 60000          # python int: no device->host sync per step (the trainer's (1,) tensor also works)

    # =========================== leg 1: device-resident inputs (`value`, weak scaling) ===========================
    # every rank has its own patch draws (seed 1000 + rank); the draws (96 B per step) are staged in HBM beforehand
    draws_host = make_draws(1000 + rank, total_steps)
    draws_dev = [torch.from_numpy(d).to(dev) for d in draws_host]
    graphed = None
    if not args.eager:
        graphed = GraphedTrainStep(net, crit, opt, frame_kw, bg01, N_PATCHES, PATCH, reducer=reducer, world=world,
                                   scaling='weak', rank=rank, sampler=sampler)
    eager_buf = sampler.alloc()
    eager_rays = [0]

    def run_step(draws):
        if graphed is not None:
            return graphed(None, it_val, draws=draws)
        b = sampler.select(draws[0].contiguous(), draws[1].to(torch.int64), out=eager_buf)
        eager_rays[0] += int(b['n_rays'])                     # eager mode only (host sync per step)
        return eager_step(b, it_val)

    for i in range(args.warmup):
        run_step(draws_dev[i])
    barrier()
    if world > 1:
        # the first multi-GPU run on a fresh box has shown a one-off ~45 ms stall a few replays after the first
        # barrier (NCCL / peer-mapping lazy initialisation): settle with more untimed replays of the warm-up batches
        for r in range(10):
            run_step(draws_dev[r % max(args.warmup, 1)])
        barrier()
    clocks = ClockSampler(local)
    clocks.start()
    launches0 = _lib.launch_count()
    rays0 = int(graphed.rays_seen) if graphed is not None else eager_rays[0]
    if graphed is None:
        ops.timing_start()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ev0.record()
    for i in range(args.warmup, total_steps):
        run_step(draws_dev[i])
    ev1.record()
    barrier()
    op_ms = ops.timing_stop() if graphed is None else {}
    launches = (_lib.launch_count() - launches0) if graphed is None else graphed.launches_per_step * args.steps
    clk = clocks.stop()
    ms = ev0.elapsed_time(ev1)
    rays_local = (int(graphed.rays_seen) if graphed is not None else eager_rays[0]) - rays0
    stats = torch.tensor([ms, float(rays_local)], device=dev, dtype=torch.float64)
    if world > 1:
        mx = stats.clone()
        dist.all_reduce(mx, op=dist.ReduceOp.MAX)
        sm = stats.clone()
        dist.all_reduce(sm, op=dist.ReduceOp.SUM)
        ms, rays_total = float(mx[0]), float(sm[1])
    else:
        rays_total = float(rays_local)
    value = rays_total / (ms / 1e3)

    # =========================== leg 2: end to end through the public API (`e2e`) =================
    # per step, inside the timed region: the host draws the step's patch randoms with numpy (the reference's
    # rand / choice calls, incl. its O(n) permutation), stages them and the schedule in pinned memory, H2D, one graph
    # replay (patch selection ... Adam), and reads the loss back
    rng_e2e = np.random.RandomState(2000 + rank)
    for i in range(3):
        run_step(torch.from_numpy(sampler.draw(rng_e2e)).to(dev) if graphed is None else sampler.draw(rng_e2e))
    barrier()
    rays0 = int(graphed.rays_seen) if graphed is not None else eager_rays[0]
    host_s = 0.0
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    dr = sampler.draw(rng_e2e)
    loss_host = [torch.zeros(1, pin_memory=True) for _ in range(2)]          # pinned landing slots of the loss read-back
    loss_ready = [torch.cuda.Event(), torch.cuda.Event()]
    losses_read = 0
    for i in range(args.steps):
        loss_val = run_step(dr if graphed is not None else torch.from_numpy(dr).to(dev))
        loss_host[i & 1].copy_(loss_val.detach().reshape(1), non_blocking=True)      # D2H of THIS step's loss, enqueued
        loss_ready[i & 1].record()
        # the next step's draws are made while the GPU runs this step (what a DataLoader worker does for the reference)
        t0 = time.perf_counter()
        dr = sampler.draw(rng_e2e)
        host_s += time.perf_counter() - t0
        # read-back one step behind: the host waits for step i - 1 (long finished) and enqueues step i + 1 while the
        # GPU still runs step i, so the ~0.4 ms of H2D + graph launch of a step hide under the previous one.  Every
        # step's loss reaches the host inside the timed region; the host never runs more than one step ahead.
        if i > 0:
            loss_ready[(i - 1) & 1].synchronize()
            _ = float(loss_host[(i - 1) & 1][0])
            losses_read += 1
    loss_ready[(args.steps - 1) & 1].synchronize()
    _ = float(loss_host[(args.steps - 1) & 1][0])
    losses_read += 1
    assert losses_read == args.steps
    e1.record()
    barrier()
    ms_e2e = max_over_ranks(e0.elapsed_time(e1))
    rays_e2e = (int(graphed.rays_seen) if graphed is not None else eager_rays[0]) - rays0
    if world > 1:
        t = torch.tensor([float(rays_e2e)], device=dev, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.SUM)
        rays_e2e = float(t[0])
    e2e = rays_e2e / (ms_e2e / 1e3)
    h2d_bytes = 2 * N_PATCHES * 8 + (int(cfg.non_rigid_motion_mlp.multires) + 2) * 4

    # =========================== N > 1: strong scaling ===========================
    strong = None
    if world > 1 and graphed is not None and not args.no_strong_leg:
        g_strong = GraphedTrainStep(net, crit, opt, frame_kw, bg01, N_PATCHES, PATCH, reducer=reducer, world=world,
                                    scaling='strong', rank=rank, sampler=sampler)
        shared = [torch.from_numpy(d).to(dev) for d in make_draws(3000, total_steps)]     # the SAME step on every rank
        for i in range(args.warmup):
            g_strong(None, it_val, draws=shared[i])
        barrier()
        r0 = int(g_strong.rays_seen)
        s0, s1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s0.record()
        for i in range(args.warmup, total_steps):
            g_strong(None, it_val, draws=shared[i])
        s1.record()
        barrier()
        ms_strong = max_over_ranks(s0.elapsed_time(s1))
        rays_strong = int(g_strong.rays_seen) - r0                      # every rank counts the whole step's rays
        strong = {'ms_per_step': ms_strong / args.steps, 'rays_per_sec': rays_strong / (ms_strong / 1e3),
                  'rays_per_step_total': rays_strong / args.steps, 'rays_per_step_per_gpu': rays_strong / args.steps / world,
                  'scaling': 'strong: one 6-patch step split contiguously over the ranks, rgb all_gather (12 B/ray), identical '
                             'loss on every rank, cut-B gradient exchange (sum)',
                  'launches_per_step': g_strong.launches_per_step}
        g_strong.graph = None
        opt.lr_scale = graphed.sched[graphed._nf + 1:graphed._nf + 2]

    # =========================== opt-in: the same step with empty-space skipping ====================
    # cfg.b200.cull_in_training (eps = 1e-7): samples below the foreground-likelihood threshold skip the MLPs
    # inside the step; gradients agree with the dense step to < 1e-6 relative
    # (tests/test_gpu_tc_train.py::test_training_with_empty_space_skipping).  Reported next to `value`, never as it.
    skip_result = None
    if graphed is not None and args.precision == 'bf16' and not args.no_skip_leg:
        cfg.b200.cull_mask_eps, cfg.b200.cull_in_training = 1e-7, True
        try:
            g_skip = GraphedTrainStep(net, crit, opt, frame_kw, bg01, N_PATCHES, PATCH, reducer=reducer, world=world,
                                      scaling='weak', rank=rank, sampler=sampler)
            for i in range(args.warmup):
                g_skip(None, it_val, draws=draws_dev[i])
            barrier()
            r0 = int(g_skip.rays_seen)
            s0, s1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            s0.record()
            for i in range(args.warmup, total_steps):
                g_skip(None, it_val, draws=draws_dev[i])
            s1.record()
            barrier()
            ms_skip = max_over_ranks(s0.elapsed_time(s1))
            kept = int(net.last_cull_count) if net.last_cull_count is not None else None
            skip_result = {'cull_mask_eps': 1e-7, 'ms_per_step': ms_skip / args.steps, 'rays_per_sec': rays_total / (ms_skip / 1e3),
                           'samples_kept_last_step_this_rank': kept,
                           'samples_per_step_this_rank': N_PATCHES * PATCH * PATCH * N_SAMPLES,
                           'gradient_rel_l2_vs_dense': '< 1e-6 (tests/test_gpu_tc_train.py::test_training_with_empty_space_skipping)'}
            g_skip.graph = None
        finally:
            cfg.b200.cull_mask_eps, cfg.b200.cull_in_training = 0.0, False
        opt.lr_scale = graphed.sched[graphed._nf + 1:graphed._nf + 2]

    # =========================== rooflines from the per-op CUDA-event times ========================
    # Kernels inside a replayed graph cannot be bracketed by events; the same steps are therefore run once
    # more EAGERLY (same kernels, same inputs, per-op CUDA events on the launching stream).
    roof_steps = args.steps
    if graphed is not None:
        roof_steps = min(args.steps, 5)
        batches = [sampler.select(d[0].contiguous(), d[1].to(torch.int64)) for d in draws_dev[:args.warmup + roof_steps]]
        for i in range(2):
            eager_step(batches[i], it_val)
        barrier()
        ops.timing_start()
        for i in range(args.warmup, args.warmup + roof_steps):
            eager_step(batches[i], it_val)
        op_ms = ops.timing_stop()
        rays_local = sum(int(b['n_rays']) for b in batches[args.warmup:args.warmup + roof_steps])
    samples_local = rays_local * N_SAMPLES
    sfx = '_' + args.precision
    fld_ms = op_ms.get('field_fwd' + sfx, (0, 0))[0] + op_ms.get('field_bwd' + sfx, (0, 0))[0]
    achieved_tf = FLOPS_FWD_BWD * samples_local / (fld_ms * 1e-3) / 1e12 if fld_ms > 0 else 0.0
    kname = ('canonical field fwd+bwd: field_tc2_kernel<store> + field_tc2_bwd_kernel x2 (cta_group::2 CTA pairs) + tile_wgrad_kernel (tcgen05, bf16)'
             if args.precision == 'bf16' else 'canonical field fwd+bwd (fp32 SIMT sgemm_kernel chain)')
    tr = (traffic or {}).get('train_field_group')
    step_samples = N_PATCHES * PATCH * PATCH * N_SAMPLES
    field_traffic = tr['dram_bytes'] * step_samples / tr['samples'] if (tr and args.precision == 'bf16') else None
    roofline = {'kernel': kname,
                'bound': 'tensor', 'achieved': achieved_tf, 'peak': peaks['tf_sustained'], 'unit': 'TFLOP/s',
                'frac': achieved_tf / peaks['tf_sustained'],
                'traffic': field_traffic, 'traffic_source': tr['source'] if tr else None,
                'peak_source': peaks['source'] + ' (sustained bf16)',
                'share_of_step': (fld_ms / roof_steps) / max(ms / args.steps, 1e-9),
                'timed': 'per-op CUDA events' + (' in an eager re-run of the same steps (the timed region replays a CUDA graph)'
                                                 if graphed is not None else ' inside the timed region')}
    if field_traffic and fld_ms > 0:
        # the same kernel group seen from the memory side (stored bf16 activations)
        gbs = field_traffic * (samples_local / roof_steps / step_samples) / (fld_ms / roof_steps * 1e-3) / 1e9
        roofline['hbm_view'] = {'bound': 'hbm', 'achieved': gbs, 'peak': peaks['hbm_gbs'], 'unit': 'GB/s',
                                'frac': gbs / peaks['hbm_gbs'], 'note': 'measured DRAM traffic (ncu) / measured time'}
    op_table = {k: {'ms_total': round(v[0], 3), 'calls': v[1]} for k, v in sorted(op_ms.items())}

    result = {
        'metric': 'train_step_rays_per_sec', 'value': value, 'unit': 'rays/s', 'n_gpus': world, 'steps': args.steps,
        'warmup': args.warmup, 'ms_per_step': ms / args.steps, 'higher_is_better': True, 'scaling': 'weak',
        'vs_baseline': None, 'dtype': 'bf16' if args.precision == 'bf16' else 'f32', 'data': 'synthetic',
        'config': workload_config(world),
        'samples_per_sec': value * N_SAMPLES,
        'e2e': {'value': e2e, 'unit': 'rays/s', 'h2d_bytes_per_step': h2d_bytes, 'd2h_bytes_per_step': 4,
                'ms_per_step': ms_e2e / args.steps, 'host_draw_ms_per_step': 1e3 * host_s / args.steps,
                'note': 'frame (rays, masks, image) resident in HBM like the network weights; per step the host draws the '
                        "reference's 2 randoms per patch (numpy, in the timed region) -> pinned -> H2D -> graph replay "
                        '(device patch selection ... Adam) -> loss D2H into pinned memory, read by the host one step behind'},
        'strong': strong,
        'step_mode': 'eager' if graphed is None else 'cuda_graph (training.GraphedTrainStep, device patch selection inside)',
        'train_empty_space_skipping': skip_result,
        'gpu_launches': int(launches), 'clocks': clk, 'roofline': roofline, 'ops_ms': op_table,
        'valid_rays_in_frame': int(n_valid),
    }

    # =========================== render leg (forward, bf16 tcgen05) ================================
    if not args.no_render:
        # the frame goes through the network in ONE chunk (cfg.chunk raised from the reference's 32768 rays: 180 GB of HBM
        # hold every per-sample tensor of a frame; chunking changes no value - tests/test_gpu_parity.py - only adds 8 x
        # the launches and a concatenation of the per-sample outputs)
        for _ in range(3):
            render_frame(sharded=True, one_chunk=True)
        barrier()
        ops.timing_start()
        r0, r1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        reps = 10          # frames back to back: the host enqueues frame i + 1 while frame i runs (no sync in between)
        r0.record()
        for _ in range(reps):
            render_frame(sharded=True, one_chunk=True)
        r1.record()
        barrier()
        rop = ops.timing_stop()
        rms = max_over_ranks(r0.elapsed_time(r1) / reps)
        a_sh, b_sh = shard_bounds(n_valid, rank, world)
        P = (b_sh - a_sh) * N_SAMPLES            # samples this rank's kernels processed (per-kernel rooflines)
        tc_ms = rop.get('field_fwd_bf16', (0, 1))[0] / reps
        mf_ms = rop.get('motion_field_fwd', (0, 1))[0] / reps
        tc_tf = FLOPS_FWD * P / (tc_ms * 1e-3) / 1e12 if tc_ms > 0 else 0.0
        mf_gbs = (MF_BYTES_NOW * P + MF_ONCE) / (mf_ms * 1e-3) / 1e9 if mf_ms > 0 else 0.0
        rt, mt = (traffic or {}).get('render_field'), (traffic or {}).get('render_motion_field')
        result['render'] = {
            'rays_per_sec': n_valid / (rms * 1e-3), 'samples_per_sec': n_valid * N_SAMPLES / (rms * 1e-3), 'ms_per_frame': rms,
            'rays': int(n_valid), 'chunk': 'whole frame in one chunk', 'scaling': 'strong (one frame, rays sharded over ranks + all_gather of 20 B/ray)', 'ops_ms_per_frame': {k: round(v[0] / reps, 3) for k, v in sorted(rop.items())},
            'roofline_mlp': {'kernel': 'field_ts2_kernel (tcgen05 cta_group::2, TS form: hidden activations in tensor memory; fused PE + both MLPs on CTA pairs)', 'bound': 'tensor', 'achieved': tc_tf,
                             'peak': peaks['tf'], 'unit': 'TFLOP/s', 'frac': tc_tf / peaks['tf'],
                             'traffic': rt['dram_bytes'] if rt else None, 'traffic_source': rt['source'] if rt else None,
                             'peak_source': peaks['source'] + ' (burst bf16)'},
            'roofline_motion_field': {'kernel': 'motion_field_fwd_kernel<false> (weights output omitted: 28 B/sample); instruction-issue bound (ncu: issue slots 80 % busy), not HBM bound - see DESIGN.md 4.3',
                                      'bound': 'hbm', 'achieved': mf_gbs, 'peak': peaks['hbm_gbs'], 'unit': 'GB/s',
                                      'frac': mf_gbs / peaks['hbm_gbs'],
                                      'traffic': mt['dram_bytes'] if mt else None, 'traffic_source': mt['source'] if mt else None,
                                      'peak_source': peaks['source']},
        }

        # opt-in empty-space skipping (cfg.b200.cull_mask_eps = 1e-7: outputs within 2.6e-5 of the dense
        # render) with the frame in one chunk; reported next to the dense leg, never instead of it
        dense_ref = render_frame(sharded=True, one_chunk=True)
        for _ in range(2):
            culled = render_frame(sharded=True, cull_eps=1e-7, one_chunk=True)
        barrier()
        c0, c1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        c0.record()
        for _ in range(reps):
            culled = render_frame(sharded=True, cull_eps=1e-7, one_chunk=True)
        c1.record()
        barrier()
        cms = max_over_ranks(c0.elapsed_time(c1) / reps)
        result['render']['empty_space_skipping'] = {
            'cull_mask_eps': 1e-7, 'ms_per_frame': cms, 'rays_per_sec': n_valid / (cms * 1e-3),
            'samples_skipped_frac': float(net.last_cull_fraction),
            'max_abs_rgb_diff_vs_dense': float((culled['rgb'] - dense_ref['rgb']).abs().max()),
            'max_abs_alpha_diff_vs_dense': float((culled['alpha'] - dense_ref['alpha']).abs().max()),
            'bound': 2 * N_SAMPLES * 1e-7}

    # =========================== baselines (rank 0, N = 1 only) =================================
    if world == 1 and not args.no_eager_gpu_baseline:
        # SURVEY 8d "second comparison column": the oracle port (plain PyTorch ops, fp32, no kernel of this repo) running
        # the SAME whole step eagerly on this GPU
        try:
            graphed_graph = graphed.graph if graphed is not None else None
            torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32 = False, False
            rays_g, sec_g = oracle_train_steps(f'cuda:{local}', steps=5, warmup=2, patch=PATCH)
            result['eager_gpu_baseline'] = {
                'value': rays_g / sec_g, 'unit': 'rays/s', 'ms_per_step': 1e3 * sec_g / 5, 'kind': 'port',
                'what': 'oracle port of the reference step (torch fp32 eager on cuda, TF32 off, host numpy patch selection like the '
                        'reference Dataset) on this B200: 5 steps after 2 warm-up, wall clock with synchronize'}
        except Exception as e:      # the baseline must never take the bench line down
            result['eager_gpu_baseline'] = {'error': repr(e)[:300]}
        finally:
            torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32 = True, False
    if world == 1 and not args.no_cpu_baseline:
        cores = os.cpu_count() or 1
        cpu_patch, cpu_steps = 16, 2
        rays_c, sec_c = oracle_train_steps('cpu', steps=cpu_steps, warmup=1, patch=cpu_patch, threads=cores)
        result['cpu_baseline'] = {'value': rays_c / sec_c, 'unit': 'rays/s', 'cores': cores, 'kind': 'port',
                                  'sample': f'{cpu_steps} whole oracle train steps (patch selection + fwd + MSE + LPIPS + bwd + Adam) '
                                            f'of {N_PATCHES} patches x {cpu_patch}x{cpu_patch} rays x {N_SAMPLES} samples after 1 warm-up '
                                            f'(bounded sample of the {PATCH}x{PATCH} workload; `--impl reference` runs the full size)'}
    if rank == 0:
        print(json.dumps(result), flush=True)
    if world > 1:
        # Tear down in a fixed order: the captured graph holds NCCL kernels, and destroying the process
        # group while it is alive hung a 2-GPU run at exit.  Drop the graph, drain the device, meet at a
        # barrier, then leave without running the communicator destructors (the result is already out).
        if graphed is not None:
            graphed.graph = None
        torch.cuda.synchronize()
        dist.barrier()
        torch.cuda.synchronize()
        sys.stdout.flush()
        sys.stderr.flush()
        os._exit(0)


if __name__ == '__main__':
    main()
