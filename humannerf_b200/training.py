"""Loss and train step around the rendering hot path (reference
`core/train/trainers/human_nerf/trainer.py:28-37, 97-175, 194-220`, and the Adam groups of
`core/train/optimizers/human_nerf/optimizer.py:12-43`).

Patch assembly uses the CUDA row scatter (ops.scatter_rows); the perceptual term is the LPIPS(VGG16)
architecture on cuDNN with one fused head kernel per stage.  The reference's LPIPS needs ImageNet VGG16 weights
that cannot be downloaded offline, so `PerceptualLoss(seed)` is the reference's own
`LPIPS(net='vgg', pnet_rand=True)` under that seed - pinned against the reference module by
tests/golden/train_step_cfg2.npz; parity against the ImageNet weights stays unpinned."""
from __future__ import annotations

from typing import Dict, Optional

import torch
import torch.nn as nn

from humannerf_b200 import ops
from humannerf_b200.config import get_cfg


def patch_scatter_index(patch_masks: torch.Tensor) -> torch.Tensor:
    """Flat (N*P*P) destination row of every ray, in ray order: patch i's rays fill its True
    pixels in row-major order (trainer.py:35 `patch_imgs[i, patch_masks[i]] = rgbs[div[i]:div[i+1]]`)."""
    return torch.nonzero(patch_masks.reshape(-1), as_tuple=False)[:, 0]


def unpack_imgs(rgbs: torch.Tensor, patch_masks: torch.Tensor, bgcolor01: torch.Tensor,
                targets: torch.Tensor, div_indices=None, index: Optional[torch.Tensor] = None) -> torch.Tensor:
    """(N_patch, P, P, 3) patches pre-filled with the background colour (trainer.py:28-37)."""
    base = bgcolor01.to(rgbs).expand(targets.shape).contiguous()
    if index is None:
        index = patch_scatter_index(patch_masks)
    return ops.scatter_rows(rgbs, index, base)


class PerceptualLoss(nn.Module):
    """LPIPS v0.1 with a VGG16 trunk (third_parties/lpips/lpips.py:23-129): scale -> five VGG16 stages
    -> unit-normalise channels (third_parties/lpips/__init__.py:40-42) -> squared difference -> 1x1 non-negative
    'lin' -> spatial mean -> sum.

    Weights.  The reference builds `LPIPS(net='vgg')` (trainer.py:70), whose trunk is torchvision's ImageNet VGG16 -
    a download, unavailable offline.  `PerceptualLoss(seed)` builds the SAME trunk the reference builds with
    `LPIPS(net='vgg', pnet_rand=True)` under `torch.manual_seed(seed)` (torchvision's constructor consumes the
    generator identically: pinned bit for bit by tests/golden/train_step_cfg2.npz), and `load_reference_weights`
    takes real weights when they exist: the trunk from a torchvision VGG16 state_dict and the five 1x1 `lin`
    layers from the reference's vendored `third_parties/lpips/weights/v0.1/vgg.pth`."""

    def __init__(self, seed: int = 0):
        super().__init__()
        import torchvision
        with torch.random.fork_rng():
            torch.manual_seed(seed)
            feats = torchvision.models.vgg16(weights=None).features
            cuts = (0, 4, 9, 16, 23, 30)
            self.stages = nn.ModuleList([nn.Sequential(*[feats[i] for i in range(a, b)]) for a, b in zip(cuts[:-1], cuts[1:])])
            self.lins = nn.ModuleList([nn.Conv2d(c, 1, 1, bias=False) for c in (64, 128, 256, 512, 512)])
            for l in self.lins:
                l.weight.data.uniform_(0.0, 2.0 / l.weight.shape[1])
        self.register_buffer('shift', torch.tensor([-.030, -.088, -.188])[None, :, None, None])
        self.register_buffer('scale', torch.tensor([.458, .448, .450])[None, :, None, None])
        for p in self.parameters():
            p.requires_grad_(False)
        # the patches arrive as NHWC images (a channels_last view after permute): NHWC filters spare cuDNN ten
        # 2.4 M-element filter conversions per step
        self.stages.to(memory_format=torch.channels_last)
        self.eval()

    def load_reference_weights(self, lins=None, trunk=None) -> 'PerceptualLoss':
        """`lins`: the reference LPIPS module's state_dict (or its vgg.pth: keys `lin{k}.model.1.weight`), or a list of
        five arrays; `trunk`: a torchvision VGG16 `features` state_dict (keys `{i}.weight` / `{i}.bias`) or the
        reference LPIPS state_dict (keys `net.slice{s}.{i}.*`)."""
        with torch.no_grad():
            if lins is not None:
                if isinstance(lins, dict):
                    lins = [lins[f'lin{k}.model.1.weight'] for k in range(5)]
                for l, w in zip(self.lins, lins):
                    l.weight.copy_(torch.as_tensor(w, dtype=torch.float32).reshape(l.weight.shape))
            if trunk is not None:
                flat = {}
                for k, v in trunk.items():
                    parts = k.split('.')
                    if parts[0] == 'net' and parts[1].startswith('slice'):
                        flat[f'{parts[2]}.{parts[3]}'] = v
                    elif parts[0] == 'features':
                        flat[f'{parts[1]}.{parts[2]}'] = v
                    elif len(parts) == 2 and parts[0].isdigit():
                        flat[k] = v
                cuts = (0, 4, 9, 16, 23, 30)
                for stage, a in zip(self.stages, cuts[:-1]):
                    for j, m in enumerate(stage):
                        if isinstance(m, nn.Conv2d):
                            m.weight.copy_(flat[f'{a + j}.weight'])
                            m.bias.copy_(flat[f'{a + j}.bias'])
        return self

    def forward(self, x: torch.Tensor, y: torch.Tensor) -> torch.Tensor:
        # prediction and target go through the trunk as ONE batch (half the launches of two passes; every op
        # of the trunk is per image, so the values are those of lpips.py:78-92)
        n = x.shape[0]
        f = (torch.cat([x, y], 0) - self.shift) / self.scale
        total = 0.0
        if f.is_cuda:
            # normalise -> squared difference -> lin -> spatial mean: one kernel per stage (hn_lpips_head_fwd / _bwd)
            acc = None
            for stage, lin in zip(self.stages, self.lins):
                f = stage(f)
                acc = ops.lpips_head(f, lin.weight, acc)
            return acc.view(n, 1, 1, 1)
        for stage, lin in zip(self.stages, self.lins):       # CPU tensors: host-logic tests and the oracle arm
            f = stage(f)
            nf = f / (torch.sqrt(torch.sum(f ** 2, dim=1, keepdim=True) + 1e-10) + 1e-10)
            total = total + lin((nf[:n] - nf[n:]) ** 2).mean([2, 3], keepdim=True)
        return total


class RenderLoss(nn.Module):
    """0.2 * MSE + 1.0 * LPIPS over the re-assembled patches (cfg.train.lossweights)."""

    def __init__(self, perceptual: Optional[nn.Module] = None):
        super().__init__()
        self.perceptual = perceptual

    def forward(self, rgb, patch_masks, bgcolor01, targets, index=None) -> Dict[str, torch.Tensor]:
        """`index` (R,) = flat destination pixel of every ray (-1: padding ray); when given, `patch_masks`
        is not needed (and no device-side nonzero / host sync happens)."""
        w = get_cfg().train.lossweights
        imgs = unpack_imgs(rgb, patch_masks, bgcolor01, targets, index=index)
        losses = {}
        if w.get('mse', 0) > 0:
            losses['mse'] = w['mse'] * torch.mean((imgs - targets) ** 2)
        if w.get('l1', 0) > 0:
            losses['l1'] = w['l1'] * torch.mean(torch.abs(imgs - targets))
        if w.get('lpips', 0) > 0 and self.perceptual is not None:
            lp = self.perceptual(imgs.permute(0, 3, 1, 2) * 2. - 1., targets.permute(0, 3, 1, 2) * 2. - 1.)
            losses['lpips'] = w['lpips'] * torch.mean(lp)
        losses['total'] = sum(losses.values())
        return losses


class FusedAdam(torch.optim.Optimizer):
    """Adam over all parameter tensors in ONE kernel launch (`hn_adam_step`, csrc/adam.cu).

    Same constructor groups and update rule as the `torch.optim.Adam` the reference builds
    (optimizer.py:12-43: one group per tensor, per-name learning rates); `param_groups[i]['lr']` stays
    the knob `update_lr` turns.  Differences in mechanism, not arithmetic:

    * first and second moments live in two flat fp32 buffers (`state[p]['exp_avg' / 'exp_avg_sq']` are
      views), the kernel's tensor table is built once;
    * gradients are taken where autograd left them: each step uploads the 55 `.grad` pointers (one
      440-byte copy from pinned memory) - `zero_grad()` just drops the old tensors, so there is no
      zero-fill and no `grad += new` pass over 64 M elements;
    * step counts, learning rates, an optional global `lr_scale` and the per-tensor `active` mask are
      read from device memory, so a step captured in a CUDA graph follows the schedule on replay;
    * a tensor whose `.grad` is None or whose `active` entry is 0 is skipped, like torch skips a
      parameter without gradient.
    """

    def __init__(self, params, lr=5e-4, betas=(0.9, 0.999), eps=1e-8):
        super().__init__(params, dict(lr=lr, betas=betas, eps=eps))
        from humannerf_b200 import _lib
        self._lib = _lib
        ps = [p for g in self.param_groups for p in g['params']]
        # dense = one storage element per tensor element: row-major, or channels_last_3d for the transposed-convolution
        # weights of the volume decoder (the kernel walks storage order; gradients must share the parameter's strides)
        def dense(p):
            return p.is_contiguous() or (p.dim() == 5 and p.is_contiguous(memory_format=torch.channels_last_3d))
        if not ps or any(not p.is_cuda or p.dtype != torch.float32 or not dense(p) for p in ps):
            raise RuntimeError('FusedAdam: parameters must be dense fp32 CUDA tensors (no CPU fallback)')
        if len({(g['betas'], g['eps']) for g in self.param_groups}) != 1:
            raise RuntimeError('FusedAdam: betas / eps must be the same for every group')
        dev = ps[0].device
        self._ps = ps
        n_t = len(ps)
        offs, total = [], 0
        for p in ps:
            offs.append(total)
            total += (p.numel() + 3) // 4 * 4                # keep every view 16-byte aligned
        self._offs = offs
        self._flat_m = torch.zeros(total, device=dev)
        self._flat_v = torch.zeros(total, device=dev)
        self.step_dev = torch.zeros(n_t, device=dev)             # per-tensor step counts (floats)
        self.active = torch.ones(n_t, device=dev)                # user mask: 0 = skip this tensor
        self.lr_scale: Optional[torch.Tensor] = None             # optional (1,) device factor on every lr
        table = (_lib.AdamTensor * n_t)()
        for i, (p, o) in enumerate(zip(ps, offs)):
            n = p.numel()
            self.state[p] = {'step': self.step_dev[i:i + 1],
                             'exp_avg': torch.as_strided(self._flat_m[o:o + n], p.shape, p.stride()),
                             'exp_avg_sq': torch.as_strided(self._flat_v[o:o + n], p.shape, p.stride())}
            table[i].param, table[i].grad = p.data_ptr(), None
            table[i].exp_avg, table[i].exp_avg_sq = self._flat_m[o:o + n].data_ptr(), self._flat_v[o:o + n].data_ptr()
            table[i].numel = n
        lib = _lib.load()
        nbytes = int(lib.hn_adam_plan_bytes(table, n_t))
        self._n_chunks = int(lib.hn_adam_n_chunks(table, n_t))
        host = torch.empty(nbytes, dtype=torch.uint8)
        _lib.check(lib.hn_adam_plan(table, n_t, host.data_ptr(), nbytes), 'adam_plan')
        self._plan = host.to(dev)
        # per-step upload: [gradient pointers (int64) | has-gradient flags (float32 bit patterns in int64 slots)]
        self._step_host = torch.zeros(2 * n_t, dtype=torch.int64).pin_memory()
        self._step_dev = torch.zeros(2 * n_t, dtype=torch.int64, device=dev)
        self._has_grad = self._step_dev[n_t:].view(torch.float32)[::2]      # low halves of the int64 slots
        self._step_active = torch.ones(n_t, device=dev)
        self._keep = []
        self._lr_cache = None
        self._lr_dev = torch.zeros(n_t, device=dev)
        self._sync_lr()

    def _lrs(self):
        return [float(g['lr']) for g in self.param_groups for _ in g['params']]

    def _adopt_state(self) -> None:
        """Copy whatever `self.state` holds (e.g. tensors a checkpoint brought in) into the flat moment buffers and the
        device step counts the kernel reads, and point `state[p]` back at views of them."""
        with torch.no_grad():
            for i, p in enumerate(self._ps):
                o, n = self._offs[i], p.numel()
                views = {'step': self.step_dev[i:i + 1],
                         'exp_avg': torch.as_strided(self._flat_m[o:o + n], p.shape, p.stride()),
                         'exp_avg_sq': torch.as_strided(self._flat_v[o:o + n], p.shape, p.stride())}
                st = self.state.get(p, {})
                for k, view in views.items():
                    src = st.get(k)
                    if src is None:
                        view.zero_()
                    elif not (torch.is_tensor(src) and src.data_ptr() == view.data_ptr()):
                        src = torch.as_tensor(src, dtype=torch.float32, device=view.device)
                        view.copy_(src.reshape(view.shape) if src.numel() == view.numel() else src)
                self.state[p] = views
        self._lr_cache = None                  # param_groups came from the checkpoint: re-upload the learning rates

    def load_state_dict(self, state_dict) -> None:
        """`torch.optim.Optimizer.load_state_dict` replaces `self.state` with the checkpoint's own tensors (the
        reference trainer resumes this way, trainer.py:377); the kernel reads the flat buffers, so the loaded
        moments and step counts are copied into them.  Accepts checkpoints of `torch.optim.Adam` too."""
        super().load_state_dict(state_dict)
        self._adopt_state()

    def _sync_lr(self):
        lrs = self._lrs()
        if lrs != self._lr_cache:
            if torch.cuda.is_current_stream_capturing():
                raise RuntimeError('FusedAdam: param_groups lr changed during CUDA graph capture; use lr_scale')
            self._lr_dev.copy_(torch.tensor(lrs, dtype=torch.float32))
            self._lr_cache = lrs

    @torch.no_grad()
    def step(self, closure=None):
        loss = closure() if closure is not None else None
        n_t = len(self._ps)
        # Captured in a graph, the copy below becomes a memcpy node that re-reads the pinned buffer on every
        # replay (it is not touched again after capture).  Eagerly, a pageable staging tensor is used: that
        # copy has consumed its source when it returns, so the next step may overwrite nothing in flight.
        capturing = torch.cuda.is_current_stream_capturing()
        host = self._step_host if capturing else torch.zeros(2 * n_t, dtype=torch.int64)
        flags = host[n_t:].view(torch.float32)
        keep = []
        for i, p in enumerate(self._ps):
            g = p.grad
            if g is None:
                host[i] = 0
                flags[2 * i] = 0.0
                continue
            if g.dtype != torch.float32 or g.shape != p.shape or g.stride() != p.stride():
                g = torch.empty_like(p).copy_(g.view_as(p) if g.shape != p.shape else g)   # the parameter's strides
            keep.append(g)
            host[i] = g.data_ptr()
            flags[2 * i] = 1.0
        self._keep = keep                              # the kernel reads these buffers asynchronously
        self._step_dev.copy_(host, non_blocking=capturing)
        torch.mul(self.active, self._has_grad, out=self._step_active)
        self._sync_lr()
        self.step_dev.add_(self._step_active)
        g0 = self.param_groups[0]
        lib, ptr = self._lib.load(), self._lib.ptr
        with torch.cuda.device(self._plan.device):
            self._lib.check(lib.hn_adam_step(ptr(self._plan), self._n_chunks, ptr(self._step_dev), ptr(self._lr_dev),
                                             ptr(self.lr_scale), ptr(self.step_dev), ptr(self._step_active),
                                             float(g0['betas'][0]), float(g0['betas'][1]), float(g0['eps']),
                                             self._lib.stream_ptr()), 'adam_step')
        return loss


def _param_groups(network: nn.Module):
    cfg = get_cfg()
    custom = [k[3:] for k in cfg.train.keys() if k.startswith('lr_')]
    groups = []
    for name, p in network.named_parameters():
        if not p.requires_grad:
            continue
        lr_name = next((c for c in custom if c in name), None)
        if lr_name is not None:
            groups.append({'params': [p], 'lr': cfg.train[f'lr_{lr_name}'], 'name': lr_name})
        else:
            groups.append({'params': [p], 'name': name})
    return groups


def make_optimizer(network: nn.Module, fused: Optional[bool] = None) -> torch.optim.Optimizer:
    """Adam with one group per tensor and per-name learning rates (optimizer.py:12-43): a parameter
    whose name contains `mweight_vol_decoder` / `pose_decoder` / `non_rigid_mlp` uses cfg.train.lr_<name>.
    On a CUDA network this is the single-launch `FusedAdam` (cfg.b200.fused_adam, default on)."""
    cfg = get_cfg()
    groups = _param_groups(network)
    on_cuda = all(p.is_cuda for g in groups for p in g['params'])
    if fused is None:
        fused = bool(cfg.get('b200', {}).get('fused_adam', True))
    if fused and on_cuda:
        return FusedAdam(groups, lr=cfg.train.lr, betas=(0.9, 0.999))
    return torch.optim.Adam(groups, lr=cfg.train.lr, betas=(0.9, 0.999))


def update_lr(optimizer: torch.optim.Optimizer, iter_step: int) -> None:
    """lr = base * 0.1 ** (iter / (lrate_decay * 1000))  (lr_updaters/exp_decay.py:7-16)."""
    cfg = get_cfg()
    decay = 0.1 ** (iter_step / (cfg.train.lrate_decay * 1000))
    for g in optimizer.param_groups:
        base = cfg.train[f"lr_{g['name']}"] if f"lr_{g['name']}" in cfg.train else cfg.train.lr
        g['lr'] = base * decay


# ---------------------------------------------------------------------------------------------
# the whole train step as one CUDA graph
# ---------------------------------------------------------------------------------------------
class _PinnedRing:
    """Small pinned staging buffers for per-step host values.  The host runs ahead of the GPU (a graph replay is one
    launch), so one buffer could be overwritten for step k+1 before the copy of step k has executed: a ring, each
    slot guarded by the event recorded after its last copy."""

    def __init__(self, shape, dtype, slots: int = 4):
        self.slots = [(torch.zeros(shape, dtype=dtype).pin_memory(), torch.cuda.Event()) for _ in range(slots)]
        self.i = 0

    def stage(self, fill, dst: torch.Tensor) -> None:
        h, ev = self.slots[self.i]
        self.i = (self.i + 1) % len(self.slots)
        ev.synchronize()                       # no-op unless the GPU is more than `slots` steps behind
        fill(h)
        dst.copy_(h, non_blocking=True)
        ev.record()


class GraphedTrainStep:
    """forward + loss + backward (+ gradient all-reduce) + Adam of `trainer.py:194-220` captured once in
    a CUDA graph and replayed per step: an eager step is ~1350 kernel launches and is bound by the
    host's launch rate (19 ms of host time for ~11 ms of GPU work); a replay is one launch.

    What makes the step capturable:
    * fixed shapes - batches come padded to `n_patches * patch_size**2` rays (data.make_train_batch
      `pad_rays=True`): padding rays have near = far = 0 and scatter_index = -1, render something that
      is never looked at and receive exactly zero gradient;
    * no host<->device traffic inside - the step-dependent scalars (hann window of the non-rigid
      encoding, condition-code gate, learning-rate decay) live in one small device tensor that
      `__call__` refreshes from pinned host memory before the replay; the bf16 weight image is
      re-packed by a kernel (ops.PackedField); Adam is the one-launch FusedAdam with device-side step
      counts; the rigid inverse of the motion bases is analytic (no LU `info` check);
    * inputs are copied into static device buffers before the replay (`load`).

    The loss returned by `__call__` is a static device scalar (valid until the next call).
    """

    def __init__(self, net: nn.Module, crit: nn.Module, opt: 'FusedAdam', frame_kw: Dict[str, torch.Tensor],
                 bgcolor01: torch.Tensor, n_patches: int, patch_size: int, reducer=None, world: int = 1,
                 rays_dtype: torch.dtype = torch.float64, scaling: str = 'weak', rank: int = 0, sampler=None):
        """`scaling` (world > 1): 'weak' - every rank steps on its OWN batch, gradients averaged; 'strong' - every
        rank loads the SAME batch and renders the contiguous slice `shard_bounds(cap, rank, world)` of its rays
        (SURVEY 8e cfg3), the rgb rows are all-gathered, every rank evaluates the identical loss, gradients summed.
        `sampler` (data.DevicePatchSampler): patch selection becomes the first kernel of the captured step - the
        per-step input is then `draws` (the reference's two host randoms per patch, 16 bytes each) instead of a batch."""
        if not isinstance(opt, FusedAdam):
            raise RuntimeError('GraphedTrainStep needs the FusedAdam optimizer (device-side schedule)')
        cfg = get_cfg()
        if cfg.perturb > 0 and str(cfg.get('b200', {}).get('jitter_rng', 'reference')) != 'device':
            raise RuntimeError("GraphedTrainStep: set cfg.b200.jitter_rng = 'device' (the reference draws the "
                               "stratified jitter on the CPU generator, which cannot be captured in a CUDA graph)")
        self.net, self.crit, self.opt, self.frame_kw, self.bg01 = net, crit, opt, frame_kw, bgcolor01
        self.reducer, self.world, self.rank, self.scaling = reducer, world, rank, scaling
        if scaling not in ('weak', 'strong'):
            raise ValueError(f'GraphedTrainStep: scaling = {scaling!r}')
        dev = bgcolor01.device
        self.cap = n_patches * patch_size * patch_size
        self.sampler = sampler
        if sampler is not None:
            if sampler.cap != self.cap:
                raise RuntimeError('GraphedTrainStep: the sampler was built for another patch geometry')
            rays_dtype = torch.float32                              # the sampler gathers the frame's fp32 rays
            self.draws = torch.zeros(2, n_patches, dtype=torch.float64, device=dev)
            self._draw_ring = _PinnedRing((2, n_patches), torch.float64)
            self._sampler_out = sampler.alloc()
        self.rays = torch.zeros(3, self.cap, 3, dtype=rays_dtype, device=dev)
        self.near = torch.zeros(self.cap, 1, device=dev)
        self.far = torch.zeros(self.cap, 1, device=dev)
        self.index = torch.full((self.cap,), -1, dtype=torch.int64, device=dev)
        self.targets = torch.zeros(n_patches, patch_size, patch_size, 3, device=dev)
        nf = int(cfg.non_rigid_motion_mlp.multires)
        self._nf = nf
        self.sched = torch.zeros(nf + 2, device=dev)                  # [hann (nf) | cond gate | lr scale]
        self._sched_ring = _PinnedRing((nf + 2,), torch.float32)
        self.loaded = torch.cuda.Event()
        if sampler is not None:
            self._sampler_out.update(rays=self.rays, near=self.near, far=self.far, scatter_index=self.index,
                                     target_patches=self.targets)
        opt.lr_scale = self.sched[nf + 1:nf + 2]
        self._base_lrs = None
        self.graph: Optional[torch.cuda.CUDAGraph] = None
        self.loss: Optional[torch.Tensor] = None
        self.rays_seen = torch.zeros(1, dtype=torch.int64, device=dev)      # real (non-padding) rays of all steps so far
        self.launches_per_step = 0
        self._iter_cfg = None

    # -- per-step host work ---------------------------------------------------------------
    def _set_schedule(self, iter_step: int) -> None:
        from humannerf_b200.embedders.hannw_fourier import hann_window_weights
        cfg = get_cfg()
        nr = cfg.non_rigid_motion_mlp
        def fill(h):
            h[:self._nf] = hann_window_weights(iter_step, self._nf, nr.kick_in_iter, nr.full_band_iter)
            h[self._nf] = 0.0 if iter_step < nr.kick_in_iter else 1.0
            h[self._nf + 1] = 0.1 ** (iter_step / (cfg.train.lrate_decay * 1000))      # exp_decay.py:7-16
        self._sched_ring.stage(fill, self.sched)
        pose_on = (not cfg.pose_decoder_off) and iter_step >= cfg.pose_decoder.get('kick_in_iter', 0)
        if self._iter_cfg is None:
            self._iter_cfg = pose_on
        elif self._iter_cfg != pose_on:
            self.graph = None                   # the pose decoder switched on: the captured graph is stale
            self._iter_cfg = pose_on

    def load(self, batch: Dict[str, torch.Tensor]) -> None:
        """Copy a padded batch (host pinned or device tensors) into the static input buffers.  The copies are
        asynchronous: a caller that REUSES pinned host buffers must not overwrite them before `self.loaded` (the
        event recorded after the copies) has completed - `loaded.synchronize()`; fresh buffers per step need nothing."""
        if batch['near'].shape[0] != self.cap:
            raise RuntimeError(f'GraphedTrainStep: batch has {batch["near"].shape[0]} rays, expected the padded '
                               f'{self.cap} (data.make_train_batch(pad_rays=True))')
        self.rays.copy_(batch['rays'], non_blocking=True)
        self.near.copy_(batch['near'], non_blocking=True)
        self.far.copy_(batch['far'], non_blocking=True)
        self.index.copy_(batch['scatter_index'], non_blocking=True)
        self.targets.copy_(batch['target_patches'], non_blocking=True)
        self.loaded.record()

    # -- the step -----------------------------------------------------------------------
    def _body(self, iter_step: int) -> torch.Tensor:
        self.opt.zero_grad(set_to_none=True)
        if self.sampler is not None:
            # a3 on the device: this step's patches from the two host randoms per patch, straight into the static buffers
            self.sampler.select(self.draws[0], self.draws[1].to(torch.int64), out=self._sampler_out)
        self.rays_seen += self._sampler_out['n_rays'] if self.sampler is not None else (self.index >= 0).sum()
        sched = dict(b200_hann_w=self.sched[:self._nf], b200_cond_gate=self.sched[self._nf:self._nf + 1])
        if self.world > 1 and self.scaling == 'strong':
            from humannerf_b200.parallel import GatherRowsWithGrad, shard_bounds
            a, b = shard_bounds(self.cap, self.rank, self.world)
            out = self.net(rays=self.rays[:, a:b], near=self.near[a:b], far=self.far[a:b], iter_val=iter_step,
                           **sched, **self.frame_kw)
            rgb = GatherRowsWithGrad.apply(out['rgb'], self.cap)
            losses = self.crit(rgb, None, self.bg01, self.targets, index=self.index)
            loss = losses['total']                  # each rank back-propagates into its own rows only: the SUM is the gradient
        else:
            out = self.net(rays=self.rays, near=self.near, far=self.far, iter_val=iter_step, **sched, **self.frame_kw)
            losses = self.crit(out['rgb'], None, self.bg01, self.targets, index=self.index)
            loss = losses['total'] / self.world
        if self.world > 1 and self.reducer is not None:
            self.reducer.set_boundary(**self.net.boundary_tensors)
            self.reducer.exchange_and_finish(loss)
        else:
            loss.backward()
        self.opt.step()
        return losses['total'].detach()

    def _capture(self, iter_step: int) -> None:
        from humannerf_b200 import _lib
        # learning rates: the graph multiplies the BASE rates by the device-side decay factor
        cfg = get_cfg()
        for g in self.opt.param_groups:
            g['lr'] = cfg.train[f"lr_{g['name']}"] if f"lr_{g['name']}" in cfg.train else cfg.train.lr
        # the warm-up iterations below are real optimizer steps on the loaded batch; undo them afterwards
        # so that capture has no side effect on the training state
        # Autograd caches one AccumulateGrad node per parameter for as long as ANY graph that uses the parameter is alive,
        # and runs it on the stream it was created on.  If an earlier eager step ran on the legacy default stream and
        # something still references its graph (the network keeps the last boundary tensors, a reducer too), the
        # captured backward would have to synchronise with that stream - illegal during capture.  Drop what we can reach.
        for holder in (self.net, getattr(self.net, 'module', None)):
            if holder is not None and hasattr(holder, 'boundary_tensors'):
                holder.boundary_tensors = {}
                holder.motion_weights_vol = None
        if self.reducer is not None:
            self.reducer.boundary = {}
        self.opt.zero_grad(set_to_none=True)
        snap_p = [p.detach().clone() for p in self.opt._ps]
        snap = (self.opt._flat_m.clone(), self.opt._flat_v.clone(), self.opt.step_dev.clone())
        side = torch.cuda.Stream()
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side):
            for _ in range(3):
                self._body(iter_step)
        torch.cuda.current_stream().wait_stream(side)
        torch.cuda.synchronize()
        graph = torch.cuda.CUDAGraph()
        n0 = _lib.launch_count()
        with torch.cuda.graph(graph):
            self.loss = self._body(iter_step)
        self.launches_per_step = _lib.launch_count() - n0
        with torch.no_grad():
            for p, q in zip(self.opt._ps, snap_p):
                p.copy_(q)
            self.opt._flat_m.copy_(snap[0])
            self.opt._flat_v.copy_(snap[1])
            self.opt.step_dev.copy_(snap[2])
        self.graph = graph

    def __call__(self, batch: Optional[Dict[str, torch.Tensor]], iter_step: int, draws=None) -> torch.Tensor:
        """`batch`: a padded batch (host pinned or device tensors), or None to reuse the loaded one; with a sampler
        pass `draws` = `sampler.draw(rng)` instead ((2, n_patches) float64: u and k per patch)."""
        self._set_schedule(int(iter_step))
        if draws is not None:
            if self.sampler is None:
                raise RuntimeError('GraphedTrainStep: `draws` needs a DevicePatchSampler')
            if torch.is_tensor(draws) and draws.is_cuda:
                self.draws.copy_(draws)                                   # already resident in HBM
            else:
                src = torch.as_tensor(draws, dtype=torch.float64)
                self._draw_ring.stage(lambda h: h.copy_(src), self.draws)
        elif batch is not None:
            self.load(batch)
        if self.graph is None:
            self._capture(int(iter_step))
        self.graph.replay()
        return self.loss
