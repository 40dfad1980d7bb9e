#!/usr/bin/env python
"""Benchmark of the HumanNeRF per-ray rendering hot path on B200.

  python bench.py --gpus N --steps K --warmup W            (N>1: launched with torchrun)
  python bench.py --impl reference ...                      (CPU arm: the oracle port of the reference)

Workload (BASELINE.json configs[1]/[2]): one training step = 6 patches of 32x32 rays of a synthetic
512x512 frame (<= 6144 rays x 128 samples), forward + loss (0.2 MSE + 1.0 LPIPS-VGG16 with seeded
random VGG weights, see humannerf_b200/training.py) + backward + Adam, ZJU-387 adventure.yaml
network shapes, fixture ("dense") weights, synthetic 24-joint pose; MLPs on the bf16 tcgen05 kernels
(--precision fp32 selects the SIMT parity path).  The step is one CUDA-graph replay
(training.GraphedTrainStep; --eager runs it op by op).  With N GPUs every rank renders its own 6 patches
of the same frame (weak scaling; per-GPU work fixed) and the gradients are exchanged with one bucketed
all-reduce at the replicated-input boundary (parallel.CutBReducer), captured inside the graph.

`value` = rays/s of the dense step with inputs resident in HBM; `e2e` = the same from pinned host batches
incl. the loss read-back.  The JSON line also carries: the per-kernel roofline of the MLP kernel group
(per-op CUDA events of an eager re-run of the same steps, DRAM traffic from the committed ncu capture), the
step with opt-in empty-space skipping, the forward-render leg (512x512 frame, rays sharded over the ranks +
gather; dense and with skipping) and the CPU baseline (oracle port on all host cores, bounded sample).
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
# DRAM bytes (ncu, read + write) of the bf16 field kernels for one 786 432-sample step; scales with the samples
FIELD_TRAFFIC = 19.46e9 if os.environ.get('HN_TC_RELU_BITS', '1') != '0' else 23.38e9


def load_peaks():
    p = os.path.join(ROOT, 'MEASURED_PEAKS.json')
    if os.path.isfile(p):
        with open(p) as f:
            d = json.load(f)
        return {'hbm_gbs': d['hbm_gbs'], 'tf': d['bf16_tflops'], 'tf_sustained': d.get('bf16_tflops_sustained', d['bf16_tflops']),
                'source': 'measured'}
    return {'hbm_gbs': 6650.0, 'tf': 1590.0, 'tf_sustained': 1400.0, 'source': 'fallback'}


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


def reference_arm(args):
    """The reference's own CPU implementation of the path = the oracle port (the Python reference
    cannot travel to the GPU box), all host threads, a bounded sample of the same train step."""
    import torch
    from humannerf_b200 import synthetic as syn
    from humannerf_b200.data import make_train_batch
    from humannerf_b200.training import PerceptualLoss
    from oracle import humannerf_oracle as orc
    rank = int(os.environ.get('RANK', '0'))
    if rank != 0:
        return
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    frame, K, E = scene_on_host(syn)
    ro, rd = syn.camera_rays_np(H, W, K, E)
    bounds = np.stack([frame['dst_bbox_min_xyz'], frame['dst_bbox_max_xyz']])
    near, far, mask, rd = orc.rays_bbox_near_far(bounds, ro, rd)
    ro_c, rd_c = ro[mask], rd[mask]
    yy, xx = np.mgrid[0:H, 0:W]
    subject = ((((xx - W / 2) / (0.16 * W)) ** 2 + ((yy - H / 2) / (0.38 * H)) ** 2) < 1.0) & mask.reshape(H, W)
    image = np.random.RandomState(1).rand(H, W, 3).astype(np.float32)
    patch = args.ref_patch
    t = lambda x: torch.from_numpy(np.ascontiguousarray(x))
    sd = {k: v.clone().requires_grad_(True) for k, v in syn.fixture_state_dict(syn.REFERENCE_PARAM_SHAPES).items()}
    fr = {k: t(v) for k, v in frame.items()}
    lp = PerceptualLoss(0)
    opt = torch.optim.Adam(list(sd.values()), lr=5e-4)
    rng = np.random.RandomState(0)
    times, rays = [], 0
    for step in range(args.warmup_ref + args.steps):
        b = make_train_batch(ro_c, rd_c, near, far, mask, subject, image, H, W, N_PATCHES, patch, rng)
        t0 = time.perf_counter()
        R = b['near'].shape[0]
        out = orc.network_forward(sd, t(b['rays']), t(b['near']), t(b['far']), fr, torch.full((1,), 60000),
                                  n_samples=N_SAMPLES, t_rand=torch.rand(R, N_SAMPLES))
        pm, tg = t(b['patch_masks']), t(b['target_patches'])
        img = orc.unpack_patches(out['rgb'], pm, fr['bgcolor'] / 255.0, b['patch_div_indices'])
        loss = 0.2 * torch.mean((img - tg) ** 2) + torch.mean(lp(img.permute(0, 3, 1, 2) * 2 - 1, tg.permute(0, 3, 1, 2) * 2 - 1))
        opt.zero_grad()
        loss.backward()
        opt.step()
        dt = time.perf_counter() - t0
        if step >= args.warmup_ref:
            times.append(dt)
            rays += R
    total = sum(times)
    val = rays / total
    sample = f'{N_PATCHES} patches of {patch}x{patch} rays x {N_SAMPLES} samples per step (bounded sample of the 32x32 workload)'
    print(json.dumps({
        'impl': 'reference', 'metric': 'train_step_rays_per_sec', 'value': val, 'unit': 'rays/s', 'n_gpus': args.gpus,
        'steps': args.steps, 'warmup': args.warmup_ref, 'ms_per_step': 1e3 * total / max(len(times), 1),
        'higher_is_better': True, 'scaling': 'weak', 'vs_baseline': None, 'dtype': 'f32', 'data': 'synthetic',
        'config': workload_config(args.gpus),
        'cpu_baseline': {'value': val, 'unit': 'rays/s', 'cores': cores, 'kind': 'port', 'sample': sample},
        'e2e': {'value': val, 'unit': 'rays/s', 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0},
    }))


def workload_config(n_gpus):
    return {'workload': f'ZJU-387 adventure.yaml train step: 6 patches x {PATCH}x{PATCH} rays x 128 samples per GPU per step '
                        '(fwd + 0.2 MSE + 1.0 LPIPS(VGG16, seeded random weights) + bwd + Adam), synthetic 512x512 frame, '
                        'fixture weights; render leg = one 512x512 frame (all bbox rays), bf16 tcgen05 MLP path',
            'rays_per_step_per_gpu': N_PATCHES * PATCH * PATCH, 'samples_per_ray': N_SAMPLES,
            'parallelism': f'dp{n_gpus} (rays/patches sharded, 1 bucketed cut-B gradient all-reduce per step)',
            'l2_policy': 'inputs change every step (new patches); activations '
                         f'{N_PATCHES * PATCH * PATCH * N_SAMPLES * 12.1e3 / 1e9:.1f} GB/step exceed the 126 MB L2'}


def main():
    global PATCH
    ap = argparse.ArgumentParser()
    ap.add_argument('--gpus', type=int, default=1)
    ap.add_argument('--steps', type=int, default=20)
    ap.add_argument('--warmup', type=int, default=5)
    ap.add_argument('--impl', default='ours', choices=['ours', 'reference'])
    ap.add_argument('--ref-patch', type=int, default=16, help='patch size of the bounded CPU sample (>= 16: the LPIPS VGG16 trunk pools four times)')
    ap.add_argument('--warmup-ref', type=int, default=1)
    ap.add_argument('--no-cpu-baseline', action='store_true')
    ap.add_argument('--no-render', action='store_true')
    ap.add_argument('--no-skip-leg', action='store_true', help='skip the extra train leg with empty-space skipping')
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
        if args.steps > 6:
            args.steps = 6
        return reference_arm(args)
    assert args.warmup >= 3, 'timing rules: at least 3 warm-up steps'

    import torch
    import torch.distributed as dist
    from humannerf_b200 import _lib, ops, synthetic as syn
    from humannerf_b200.config import make_cfg
    from humannerf_b200.data import make_train_batch
    from humannerf_b200.network import Network
    from humannerf_b200.parallel import CutBReducer, all_gather_rows, shard_bounds
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

    cfg = make_cfg()
    cfg.perturb = 1.0
    cfg.b200.jitter_rng = 'device'
    cfg.b200.heavy_outputs = False
    cfg.N_samples = N_SAMPLES
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
    o64, d64 = ops.gen_rays(H, W, K, E, dev)
    near64, far64, mask = ops.rays_bbox(bounds, o64, d64)
    o_c, d_c, near_c, far_c, pix = ops.compact_rays(mask, o64, d64, near64, far64)
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
    subject = (alpha > 0.5).reshape(H, W).cpu().numpy()
    if subject.sum() < 100:
        yy, xx = np.mgrid[0:H, 0:W]
        subject = (((xx - W / 2) / (0.16 * W)) ** 2 + ((yy - H / 2) / (0.38 * H)) ** 2) < 1.0
    mask_np = mask.cpu().numpy().astype(bool)
    subject &= mask_np.reshape(H, W)
    image_np = image.reshape(H, W, 3).cpu().numpy()
    o_np, d_np = o_c.double().cpu().numpy(), d_c.double().cpu().numpy()
    near_np, far_np = near_c[:, 0].cpu().numpy(), far_c[:, 0].cpu().numpy()

    # ---- per-step host batches (new patches every step), pinned like a DataLoader with pin_memory
    total_steps = args.warmup + args.steps
    rng = np.random.RandomState(1000 + rank)
    host_batches = []
    for _ in range(total_steps):
        b = make_train_batch(o_np, d_np, near_np, far_np, mask_np, subject, image_np, H, W, N_PATCHES, PATCH, rng,
                             pad_rays=True)
        hb = {k: torch.from_numpy(np.ascontiguousarray(v)).pin_memory() for k, v in b.items()
              if k not in ('patch_div_indices', 'n_rays', 'patch_masks')}
        hb['n_rays'] = int(b['n_rays'])
        host_batches.append(hb)
    h2d_bytes = int(np.mean([sum(v.numel() * v.element_size() for k, v in hb.items() if torch.is_tensor(v))
                             for hb in host_batches]))
    bg01 = fr['bgcolor'] / 255.0
    reducer = CutBReducer(net)

    def eager_step(batch, it):
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

    graphed = None
    if not args.eager:
        graphed = GraphedTrainStep(net, crit, opt, frame_kw, bg01, N_PATCHES, PATCH, reducer=reducer, world=world)

    def train_step(batch, it):
        # batch: pinned host tensors (e2e leg) or device tensors (`value` leg); the graphed step copies
        # them into its static input buffers, the eager step consumes device tensors directly
        if graphed is not None:
            return graphed(batch, it)
        return eager_step(batch, it)

    def to_device(hb):
        return {k: v.to(dev, non_blocking=True) for k, v in hb.items() if torch.is_tensor(v)}

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    it_val = 60000          # python int: no device->host sync per step (the trainer's (1,) tensor also works)

    # =========================== leg 1: device-resident inputs (`value`) ===========================
    dev_batches = [to_device(hb) for hb in host_batches]
    for i in range(args.warmup):
        train_step(dev_batches[i], it_val)
    barrier()
    if world > 1:
        # the first multi-GPU run on a fresh box has shown a one-off ~45 ms stall a few replays after the first
        # barrier (NCCL / peer-mapping lazy initialisation): settle with more untimed replays of the warm-up batches
        for r in range(10):
            train_step(dev_batches[r % max(args.warmup, 1)], it_val)
        barrier()
    clocks = ClockSampler(local)
    clocks.start()
    launches0 = _lib.launch_count()
    if graphed is None:
        ops.timing_start()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ev0.record()
    for i in range(args.warmup, total_steps):
        train_step(dev_batches[i], it_val)
    ev1.record()
    barrier()
    op_ms = ops.timing_stop() if graphed is None else {}
    launches = (_lib.launch_count() - launches0) if graphed is None else graphed.launches_per_step * args.steps
    clk = clocks.stop()
    ms = ev0.elapsed_time(ev1)
    rays_local = sum(hb['n_rays'] for hb in host_batches[args.warmup:])
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
    feed = (lambda hb: hb) if graphed is not None else to_device      # graphed: H2D straight into the static buffers
    for i in range(3):
        train_step(feed(host_batches[i]), it_val)
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(args.warmup, total_steps):
        loss_val = train_step(feed(host_batches[i]), it_val)
        _ = float(loss_val.item())                                     # device -> host read of the step's result
    e1.record()
    barrier()
    ms_e2e = e0.elapsed_time(e1)
    if world > 1:
        t = torch.tensor([ms_e2e], device=dev, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms_e2e = float(t[0])
    e2e = rays_total / (ms_e2e / 1e3)

    # =========================== opt-in: the same step with empty-space skipping ====================
    # cfg.b200.cull_in_training (eps = 1e-7): samples below the foreground-likelihood threshold skip the MLPs
    # inside the step; gradients agree with the dense step to < 1e-6 relative
    # (tests/test_gpu_tc_train.py::test_training_with_empty_space_skipping).  Reported next to `value`, never as it.
    skip_result = None
    if graphed is not None and args.precision == 'bf16' and not args.no_skip_leg:
        cfg.b200.cull_mask_eps, cfg.b200.cull_in_training = 1e-7, True
        try:
            g_skip = GraphedTrainStep(net, crit, opt, frame_kw, bg01, N_PATCHES, PATCH, reducer=reducer, world=world)
            for i in range(args.warmup):
                g_skip(dev_batches[i], it_val)
            barrier()
            s0, s1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            s0.record()
            for i in range(args.warmup, total_steps):
                g_skip(dev_batches[i], it_val)
            s1.record()
            barrier()
            ms_skip = s0.elapsed_time(s1)
            if world > 1:
                t = torch.tensor([ms_skip], device=dev, dtype=torch.float64)
                dist.all_reduce(t, op=dist.ReduceOp.MAX)
                ms_skip = float(t[0])
            kept = int(net.last_cull_count) if net.last_cull_count is not None else None
            skip_result = {'cull_mask_eps': 1e-7, 'ms_per_step': ms_skip / args.steps, 'rays_per_sec': rays_total / (ms_skip / 1e3),
                           'samples_kept_last_step_this_rank': kept,
                           'samples_per_step_this_rank': N_PATCHES * PATCH * PATCH * N_SAMPLES,
                           'gradient_rel_l2_vs_dense': '< 1e-6 (tests/test_gpu_tc_train.py::test_training_with_empty_space_skipping)'}
            g_skip.graph = None
        finally:
            cfg.b200.cull_mask_eps, cfg.b200.cull_in_training = 0.0, False
        if graphed is not None:
            opt.lr_scale = graphed.sched[graphed._nf + 1:graphed._nf + 2]

    # =========================== rooflines from the per-op CUDA-event times ========================
    # Kernels inside a replayed graph cannot be bracketed by events; the same steps are therefore run once
    # more EAGERLY (same kernels, same inputs, per-op CUDA events on the launching stream).
    roof_steps = args.steps
    if graphed is not None:
        roof_steps = min(args.steps, 5)
        for i in range(2):
            eager_step(dev_batches[i], it_val)
        barrier()
        ops.timing_start()
        for i in range(args.warmup, args.warmup + roof_steps):
            eager_step(dev_batches[i], it_val)
        op_ms = ops.timing_stop()
        rays_local = sum(hb['n_rays'] for hb in host_batches[args.warmup:args.warmup + roof_steps])
    samples_local = rays_local * N_SAMPLES
    sfx = '_' + args.precision
    fld_ms = op_ms.get('field_fwd' + sfx, (0, 0))[0] + op_ms.get('field_bwd' + sfx, (0, 0))[0]
    achieved_tf = FLOPS_FWD_BWD * samples_local / (fld_ms * 1e-3) / 1e12 if fld_ms > 0 else 0.0
    kname = ('canonical field fwd+bwd: field_tc2_kernel<store> + field_tc2_bwd_kernel x2 (cta_group::2 CTA pairs) + tile_wgrad_kernel (tcgen05, bf16)'
             if args.precision == 'bf16' else 'canonical field fwd+bwd (fp32 SIMT sgemm_kernel chain)')
    roofline = {'kernel': kname,
                'bound': 'tensor', 'achieved': achieved_tf, 'peak': peaks['tf_sustained'], 'unit': 'TFLOP/s',
                'frac': achieved_tf / peaks['tf_sustained'],
                # dram__bytes_read + write of the four launches of one 6144-ray step, ncu --set full
                # (profiles/r01g_ncu_full_summary.json): 5.06 (forward) + 3.48 + 1.29 (chains) + 9.62 (wgrad) GB;
                # with HN_TC_RELU_BITS=0 the chains read the activations again: 23.38 GB (r01e)
                'traffic': (FIELD_TRAFFIC * N_PATCHES * PATCH * PATCH * N_SAMPLES / 786432) if args.precision == 'bf16' else None,
                'peak_source': peaks['source'] + ' (sustained bf16)',
                'share_of_step': (fld_ms / roof_steps) / max(ms / args.steps, 1e-9),
                'timed': 'per-op CUDA events' + (' in an eager re-run of the same steps (the timed region replays a CUDA graph)'
                                                 if graphed is not None else ' inside the timed region')}
    if args.precision == 'bf16' and fld_ms > 0:
        # the same kernel group seen from the memory side: the stored-activation design moves 32 KB / sample
        gbs = FIELD_TRAFFIC * (samples_local / roof_steps / 786432) / (fld_ms / roof_steps * 1e-3) / 1e9
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
                'ms_per_step': ms_e2e / args.steps},
        'step_mode': 'eager' if graphed is None else 'cuda_graph (training.GraphedTrainStep)',
        'train_empty_space_skipping': skip_result,
        'gpu_launches': int(launches), 'clocks': clk, 'roofline': roofline, 'ops_ms': op_table,
        'valid_rays_in_frame': int(n_valid),
    }

    # =========================== render leg (forward, bf16 tcgen05) ================================
    if not args.no_render:
        for _ in range(3):
            render_frame(sharded=True)
        barrier()
        ops.timing_start()
        r0, r1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        reps = 5
        r0.record()
        for _ in range(reps):
            render_frame(sharded=True)
        r1.record()
        barrier()
        rop = ops.timing_stop()
        rms = r0.elapsed_time(r1) / reps
        if world > 1:
            t = torch.tensor([rms], device=dev, dtype=torch.float64)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            rms = float(t[0])
        a_sh, b_sh = shard_bounds(n_valid, rank, world)
        P = (b_sh - a_sh) * N_SAMPLES            # samples this rank's kernels processed (per-kernel rooflines)
        tc_ms = rop.get('field_fwd_bf16', (0, 1))[0] / reps
        mf_ms = rop.get('motion_field_fwd', (0, 1))[0] / reps
        tc_tf = FLOPS_FWD * P / (tc_ms * 1e-3) / 1e12 if tc_ms > 0 else 0.0
        mf_gbs = (MF_BYTES_NOW * P + MF_ONCE) / (mf_ms * 1e-3) / 1e9 if mf_ms > 0 else 0.0
        result['render'] = {
            'rays_per_sec': n_valid / (rms * 1e-3), 'samples_per_sec': n_valid * N_SAMPLES / (rms * 1e-3), 'ms_per_frame': rms,
            'rays': int(n_valid), 'scaling': 'strong (one frame, rays sharded over ranks + all_gather of 20 B/ray)', 'ops_ms_per_frame': {k: round(v[0] / reps, 3) for k, v in sorted(rop.items())},
            'roofline_mlp': {'kernel': 'field_tc2_kernel (tcgen05 cta_group::2 / TMEM, fused PE + MLPs on CTA pairs)', 'bound': 'tensor', 'achieved': tc_tf,
                             'peak': peaks['tf'], 'unit': 'TFLOP/s', 'frac': tc_tf / peaks['tf'],
                             'traffic': 176.0e6,       # per 32768-ray launch, ncu --set full (profiles/r01e_ncu_full_summary.json)
                             'peak_source': peaks['source'] + ' (burst bf16)'},
            'roofline_motion_field': {'kernel': 'motion_field_fwd_kernel<false> (weights output omitted: 28 B/sample); instruction-issue bound (ncu: issue slots 80 % busy), not HBM bound - see DESIGN.md 4.3',
                                      'bound': 'hbm', 'achieved': mf_gbs, 'peak': peaks['hbm_gbs'], 'unit': 'GB/s',
                                      'frac': mf_gbs / peaks['hbm_gbs'],
                                      'traffic': 33.1e6,    # per 32768-ray launch (profiles/r01e_ncu_full_summary.json)
                                      'issue_slots_busy_pct_ncu': 86.2, 'peak_source': peaks['source']},
        }

        # opt-in empty-space skipping (cfg.b200.cull_mask_eps = 1e-7: outputs within 2.6e-5 of the dense
        # render) with the frame in one chunk; reported next to the dense leg, never instead of it
        dense_ref = render_frame(sharded=True)
        for _ in range(2):
            culled = render_frame(sharded=True, cull_eps=1e-7, one_chunk=True)
        barrier()
        c0, c1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        c0.record()
        for _ in range(reps):
            culled = render_frame(sharded=True, cull_eps=1e-7, one_chunk=True)
        c1.record()
        barrier()
        cms = c0.elapsed_time(c1) / reps
        if world > 1:
            t = torch.tensor([cms], device=dev, dtype=torch.float64)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            cms = float(t[0])
        result['render']['empty_space_skipping'] = {
            'cull_mask_eps': 1e-7, 'ms_per_frame': cms, 'rays_per_sec': n_valid / (cms * 1e-3),
            'samples_skipped_frac': float(net.last_cull_fraction),
            'max_abs_rgb_diff_vs_dense': float((culled['rgb'] - dense_ref['rgb']).abs().max()),
            'max_abs_alpha_diff_vs_dense': float((culled['alpha'] - dense_ref['alpha']).abs().max()),
            'bound': 2 * N_SAMPLES * 1e-7}

    # =========================== CPU baseline (rank 0, N = 1 only) =================================
    if world == 1 and not args.no_cpu_baseline:
        from oracle import humannerf_oracle as orc
        cores = os.cpu_count() or 1
        torch.set_num_threads(cores)
        sd = {k: v.detach().cpu().clone().requires_grad_(True) for k, v in net.state_dict().items()}
        frc = {k: v.cpu() for k, v in fr.items()}
        hb = host_batches[0]
        R = min(512, hb['n_rays'])
        rays_cpu, near_cpu, far_cpu = hb['rays'][:, :R].clone(), hb['near'][:R].clone(), hb['far'][:R].clone()
        ts = []
        for rep in range(3):
            t0 = time.perf_counter()
            o = orc.network_forward(sd, rays_cpu, near_cpu, far_cpu, frc, torch.full((1,), 60000), n_samples=N_SAMPLES,
                                    t_rand=torch.rand(R, N_SAMPLES))
            (o['rgb'] ** 2).mean().backward()
            ts.append(time.perf_counter() - t0)
        best = min(ts[1:])
        result['cpu_baseline'] = {'value': R / best, 'unit': 'rays/s', 'cores': cores, 'kind': 'port',
                                  'sample': f'oracle fwd + MSE-style loss + bwd on the first {R} rays x {N_SAMPLES} samples of one step '
                                            f'(no LPIPS/Adam), best of 2 after 1 warm-up'}
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
