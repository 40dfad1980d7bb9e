"""`Network`: drop-in replacement of reference `core/nets/human_nerf/network.py` whose per-ray
path runs on the hand-written sm_100a kernels (humannerf_b200/csrc via the C ABI).

Same constructor (no arguments, reads the global cfg), same `forward` keyword arguments, same
11-key output dict, same parameter names (`mweight_vol_decoder.*`, `non_rigid_mlp.module.*`,
`cnl_mlp.module.*`, `pose_decoder.*`), so the reference's `train.py` / `run.py`, optimizer
(per-name learning rates) and checkpoints keep working when `cfg.network_module` points here.

Per forward (reference network.py:647-789):
  per frame, PyTorch (batch 1):  pose refinement, motion bases, motion-weight volume decode
  per chunk of cfg.chunk rays, CUDA kernels:
      ops.sample_z -> ops.motion_field -> ops.canonical_field -> ops.composite
There is no CPU fallback: tensors must live on a CUDA device and the shared library must load.
"""
from __future__ import annotations

import importlib
import importlib.util
import os
from typing import Dict, Optional

import torch
import torch.nn as nn

from humannerf_b200 import _lib, ops
from humannerf_b200.config import get_cfg
from humannerf_b200.embedders.hannw_fourier import hann_window_weights
from humannerf_b200.nn_util import SMPL_PARENT, MotionBasisComputer


def _load_symbol(module_name: str, symbol: str):
    """cfg.<component>.module -> python symbol.  Dotted module names resolve through the normal
    import system first (package installs), then - like the reference's `imp.load_source`
    (component_factory.py:3-36) - as a file path relative to the current directory."""
    try:
        return getattr(importlib.import_module(module_name), symbol)
    except ModuleNotFoundError:
        path = module_name.replace('.', '/') + '.py'
        if not os.path.isfile(path):
            raise
        spec = importlib.util.spec_from_file_location(module_name, path)
        mod = importlib.util.module_from_spec(spec)
        spec.loader.exec_module(mod)
        return getattr(mod, symbol)


class _SingleDeviceParallel(nn.Module):
    """Keeps the `.module.` level that `nn.DataParallel` puts into the reference's parameter names
    (network.py:68-72,115-119).  One process drives one GPU here, so nothing is replicated."""

    def __init__(self, module: nn.Module):
        super().__init__()
        self.module = module

    def forward(self, *args, **kwargs):
        return self.module(*args, **kwargs)


_UNSUPPORTED = (
    ('non_rigid_motion_model', lambda c: c.get('non_rigid_motion_model', 'mlp') != 'mlp'),
    ('multihead.head_num', lambda c: c.get('multihead', {}).get('head_num', 1) != 1),
    ('rgb_history.last_num', lambda c: c.get('rgb_history', {}).get('last_num', 0) != 0),
    ('canonical_mlp.view_dir', lambda c: bool(c.canonical_mlp.get('view_dir', False))),
    ('canonical_mlp.time_input', lambda c: bool(c.canonical_mlp.get('time_input', False))),
    ('canonical_mlp.input_encoder.enable', lambda c: bool(c.canonical_mlp.get('input_encoder', {}).get('enable', False))),
    ('canonical_mlp.condition_code_dim', lambda c: c.canonical_mlp.get('condition_code_dim', 0) != 0),
    ('canonical_mlp.pose_color', lambda c: c.canonical_mlp.get('pose_color', 'wo') != 'wo'),
    ('non_rigid_motion_mlp.time_input', lambda c: bool(c.non_rigid_motion_mlp.get('time_input', False))),
    ('non_rigid_motion_mlp.pose_input', lambda c: not c.non_rigid_motion_mlp.get('pose_input', True)),
    ('condition_code.type', lambda c: c.get('condition_code', {}).get('type', 'global') != 'global'),
    ('posevec.type', lambda c: c.get('posevec', {}).get('type', 'axis_angle') != 'axis_angle'),
)


# device index -> stream of the volume decoder branch (module level: streams are neither copied nor pickled with a Network)
_SIDE_STREAMS: Dict[int, 'torch.cuda.Stream'] = {}


class Network(nn.Module):
    def __init__(self):
        super().__init__()
        cfg = get_cfg()
        for key, bad in _UNSUPPORTED:
            if bad(cfg):
                raise NotImplementedError(
                    f'humannerf_b200.Network: config option `{key}` selects a research variant of the fork '
                    f'that the B200 hot path does not implement (SURVEY.md section 2, row 22)')

        self.motion_basis_computer = MotionBasisComputer(total_bones=cfg.total_bones)

        self.mweight_vol_decoder = _load_symbol(cfg.mweight_volume.module, 'MotionWeightVolumeDecoder')(
            embedding_size=cfg.mweight_volume.embedding_size,
            volume_size=cfg.mweight_volume.volume_size,
            total_bones=cfg.total_bones)

        self.get_non_rigid_embedder = _load_symbol(cfg.non_rigid_embedder.module, 'get_embedder')
        if cfg.non_rigid_motion_mlp.i_embed != 0 or cfg.canonical_mlp.i_embed != 0:
            raise NotImplementedError('humannerf_b200.Network: identity embedders (i_embed = -1) are not supported')
        nr_pe = 3 * 2 * cfg.non_rigid_motion_mlp.multires
        nr_cfg = cfg.non_rigid_motion_mlp
        self.non_rigid_mlp = _SingleDeviceParallel(
            _load_symbol(nr_cfg.module, 'NonRigidMotionMLP')(
                pos_embed_size=nr_pe,
                condition_code_size=nr_cfg.condition_code_size,
                mlp_width=nr_cfg.mlp_width,
                mlp_depth=nr_cfg.mlp_depth,
                skips=nr_cfg.skips,
                multihead_enable=nr_cfg.multihead.enable,
                multihead_depth=nr_cfg.multihead.head_depth,
                multihead_num=cfg.multihead.head_num,
                last_linear_scale=nr_cfg.last_linear_scale,
                mlp_depth_plus=nr_cfg.mlp_depth_plus))

        get_embedder = _load_symbol(cfg.embedder.module, 'get_embedder')
        self.pos_embed_fn, cnl_pe = get_embedder(cfg.canonical_mlp.multires, cfg.canonical_mlp.i_embed)
        self.dir_embed_fn = None
        cn_cfg = cfg.canonical_mlp
        self.cnl_mlp = _SingleDeviceParallel(
            _load_symbol(cn_cfg.module, 'CanonicalMLP')(
                input_ch=cnl_pe,
                mlp_depth=cn_cfg.mlp_depth,
                mlp_width=cn_cfg.mlp_width,
                view_dir=cn_cfg.view_dir,
                input_ch_dir=-1,
                pose_color=cn_cfg.pose_color,
                pose_ch=cn_cfg.pose_ch,
                skips=[4],
                multihead_enable=cn_cfg.multihead.enable,
                multihead_depth=cn_cfg.multihead.head_depth,
                multihead_num=cfg.multihead.head_num,
                last_linear_scale=cn_cfg.last_linear_scale,
                mlp_depth_plus=cn_cfg.mlp_depth_plus,
                rgb_dynamic_features_ch=0))

        if not cfg.pose_decoder_off:
            self.pose_decoder = _load_symbol(cfg.pose_decoder.module, 'BodyPoseRefiner')(
                embedding_size=cfg.pose_decoder.embedding_size,
                mlp_width=cfg.pose_decoder.mlp_width,
                mlp_depth=cfg.pose_decoder.mlp_depth)

        nr, cn = self.non_rigid_mlp.module, self.cnl_mlp.module
        # the fused field kernels plan exactly one skip connection per MLP (what every reference config uses);
        # more would silently read the wider Linears with the wrong K
        for name, mod, cats in (('non_rigid_motion_mlp', nr, getattr(nr, 'layers_to_cat_inputs', [])),
                                ('canonical_mlp', cn, getattr(cn, 'layers_to_cat_input', []))):
            if len(cats) > 1:
                raise NotImplementedError(
                    f'humannerf_b200.Network: `{name}.skips` with {len(cats)} skip connections is not supported by '
                    f'the fused field kernels (one skip per MLP)')
        self._spec = ops.FieldSpec(
            nr_layers=len(nr.linear_layers()), nr_width=nr_cfg.mlp_width, nr_skip=nr.skip_layer_index(),
            nr_freqs=nr_cfg.multires, cond_dim=nr_cfg.condition_code_size,
            cn_layers=len(cn.linear_layers()), cn_width=cn_cfg.mlp_width, cn_skip=cn.skip_layer_index(),
            cn_freqs=cn_cfg.multires)
        self._packed_field = ops.PackedField(self._spec)
        self._vol_cache = None
        self.last_cull_fraction = 0.0
        self.last_cull_count = None
        self.motion_weights_vol = None
        self.boundary_tensors = {}

    # -- reference API -------------------------------------------------------------------
    def deploy_mlps_to_secondary_gpus(self):
        """The reference scatters MLP chunks over secondary GPUs with nn.DataParallel
        (network.py:161-166).  Here every rank owns one GPU and shards rays instead
        (humannerf_b200/parallel.py), so there is nothing to deploy."""
        return self

    # -- helpers ------------------------------------------------------------------------
    def _field_weights(self):
        nr = self.non_rigid_mlp.module.linear_layers()
        cn = self.cnl_mlp.module.linear_layers()
        return ([l.weight for l in nr], [l.bias for l in nr], [l.weight for l in cn], [l.bias for l in cn])

    @staticmethod
    def _multiply_corrected_Rs(Rs, correct_Rs):
        total = Rs.shape[1]
        return torch.matmul(Rs.reshape(-1, 3, 3), correct_Rs.reshape(-1, 3, 3)).reshape(-1, total, 3, 3)

    def _decode_volume(self, priors):
        cfg = get_cfg()
        cache_ok = (not self.training) and (not torch.is_grad_enabled()) and \
            bool(cfg.get('b200', {}).get('cache_eval_volume', False))
        # the decoded volume depends only on the decoder's parameters and the prior (deconv_vol_decoder.py:25-33),
        # not on the pose: across eval frames it is the same tensor.  The cache key includes the parameters'
        # in-place version counters, so an optimizer step in between invalidates it.
        key = (priors.data_ptr(), priors._version, tuple(priors.shape),
               tuple(p._version for p in self.mweight_vol_decoder.parameters()),
               bool(getattr(self.mweight_vol_decoder, 'allow_tf32', False)))
        if cache_ok and self._vol_cache is not None and self._vol_cache[0] == key:
            return self._vol_cache[1]
        vol = self.mweight_vol_decoder(motion_weights_priors=priors)
        if cache_ok:
            self._vol_cache = (key, vol)
        return vol

    def forward(self, rays, dst_Rs, dst_Ts, cnl_gtfms, motion_weights_priors, dst_posevec=None,
                near=None, far=None, iter_val=1e7, **kwargs) -> Dict[str, torch.Tensor]:
        cfg = get_cfg()
        b200 = cfg.get('b200', {})
        dev = dst_Rs.device
        if dev.type != 'cuda':
            raise RuntimeError('humannerf_b200.Network runs on CUDA devices only (no CPU fallback)')
        it = float(torch.as_tensor(iter_val).reshape(-1)[0])

        dst_Rs, dst_Ts = dst_Rs[None, ...].float(), dst_Ts[None, ...].float()
        dst_posevec = dst_posevec[None, ...].float()
        cnl_gtfms = cnl_gtfms[None, ...].float()
        priors = motion_weights_priors[None, ...] if motion_weights_priors.dim() == 4 else motion_weights_priors

        # 0. the blend-weight volume depends on neither the pose nor the rays: its decoder (35 launches, ~0.3 ms) starts
        #    on a side stream and runs under the pose decoder / forward kinematics / weight re-pack below.  Autograd runs
        #    a node's backward on the stream of its forward, so the decoder's backward overlaps the pose branch's too.
        precision = _lib.PRECISION_BF16 if str(b200.get('mlp_precision', 'fp32')).lower() == 'bf16' \
            else _lib.PRECISION_FP32
        if hasattr(self.mweight_vol_decoder, 'allow_tf32'):
            tf32 = b200.get('decoder_tf32', 'auto')
            self.mweight_vol_decoder.allow_tf32 = (precision == _lib.PRECISION_BF16) if tf32 == 'auto' else bool(tf32)
        side = None
        if bool(b200.get('front_streams', True)):
            main = torch.cuda.current_stream(dev)
            side = _SIDE_STREAMS.get(dev.index)
            if side is None:
                side = _SIDE_STREAMS[dev.index] = torch.cuda.Stream(dev)
            pri = priors.float()
            side.wait_stream(main)
            with torch.cuda.stream(side):
                vol = self._decode_volume(pri)[0]

        # 1. pose refinement (network.py:667-695)
        refine = (not cfg.pose_decoder_off) and it >= cfg.pose_decoder.get('kick_in_iter', 0)
        fused_front = bool(b200.get('fused_front_end', True))
        rvec = None
        if refine:
            pd_out = self.pose_decoder(dst_posevec)
            if fused_front and 'rvec' in pd_out:
                rvec = pd_out['rvec'].reshape(-1, 3)          # Rodrigues happens inside ops.pose_fk
            else:
                fused_front = False
                dst_Rs = torch.cat([dst_Rs[:, 0:1], self._multiply_corrected_Rs(dst_Rs[:, 1:], pd_out['Rs'])], dim=1)

        # 2. condition code and hann window (network.py:697-738)
        nr_cfg = cfg.non_rigid_motion_mlp
        cond = dst_posevec[0]
        if kwargs.get('b200_hann_w') is not None:
            # schedule values supplied as DEVICE tensors (training.GraphedTrainStep): the hann window (6,)
            # and the 0/1 factor that zeroes the condition code before kick-in - no host->device copy
            hann = kwargs['b200_hann_w']
            cond = cond * kwargs['b200_cond_gate']
        else:
            if it < nr_cfg.kick_in_iter:
                cond = torch.zeros_like(cond)
            hann = hann_window_weights(it, nr_cfg.multires, nr_cfg.kick_in_iter, nr_cfg.full_band_iter).to(dev)

        # 3. motion bases and blend-weight volume (network.py:757-763)
        if fused_front:
            # Rodrigues + corrected rotations + the forward-kinematics chain in one launch (csrc/pose_fk.cu);
            # the LU inverse (the reference's torch.inverse without its host-side check) and the product
            # with the canonical transforms stay torch ops
            _, G = ops.pose_fk(rvec, dst_Rs[0], dst_Ts[0], SMPL_PARENT[:dst_Rs.shape[1]])
            inv = torch.linalg.inv_ex(G, check_errors=False).inverse
            f = torch.matmul(cnl_gtfms[0], inv)
            scale_Rs, Ts = f[None, :, :3, :3], f[None, :, :3, 3]
        else:
            scale_Rs, Ts = self.motion_basis_computer(dst_Rs, dst_Ts, cnl_gtfms)
        if side is not None:
            main.wait_stream(side)
            vol.record_stream(main)             # allocated on the side stream, consumed (and later freed) on this one
        else:
            vol = self._decode_volume(priors.float())[0]
        self.motion_weights_vol = vol
        # outputs of the replicated per-frame front end; the multi-GPU train step exchanges the
        # gradients of these instead of the 63.6 M decoder parameters (parallel.CutBReducer)
        self.boundary_tensors = {'motion_weights_vol': vol, 'motion_scale_Rs': scale_Rs, 'motion_Ts': Ts}
        J = cfg.total_bones
        vol_j = vol[:J]

        # 4. rays (network.py:771-777)
        rays_o = rays[0].reshape(-1, 3).float().contiguous()
        rays_d = rays[1].reshape(-1, 3).float().contiguous()
        lead_shape = list(rays[1].shape[:-1])
        near = near.reshape(-1).float()
        far = far.reshape(-1).float()
        bbox_min = kwargs['cnl_bbox_min_xyz'].float()
        bbox_scale = kwargs['cnl_bbox_scale_xyz'].float()
        bgcolor = kwargs.get('bgcolor')
        bgcolor = torch.zeros(3, device=dev) if bgcolor is None else bgcolor.float()
        t_rand_all = kwargs.get('t_rand')

        S = int(cfg.N_samples)
        heavy = bool(b200.get('heavy_outputs', True))
        cull_eps = float(b200.get('cull_mask_eps', 0.0))
        cull_train = bool(b200.get('cull_in_training', False))
        ignore_nr = bool(cfg.ignore_non_rigid_motions)
        nr_w, nr_b, cn_w, cn_b = self._field_weights()
        packed = None
        if precision == _lib.PRECISION_BF16:
            packed = self._packed_field.refresh(nr_w, nr_b, cn_w, cn_b, cond)

        outs = []
        n_kept = 0
        R = rays_o.shape[0]
        chunk = int(cfg.chunk)
        for a in range(0, R, chunk):
            b = min(R, a + chunk)
            o_c, d_c = rays_o[a:b], rays_d[a:b]
            t_rand = None
            if cfg.perturb > 0.:
                if t_rand_all is not None:
                    t_rand = t_rand_all[a:b]
                elif str(b200.get('jitter_rng', 'reference')) == 'device':
                    t_rand = torch.rand(b - a, S, device=dev)
                else:   # the reference draws on the CPU default generator (network.py:468)
                    t_rand = torch.rand(b - a, S).to(dev)
            z = ops.sample_z(near[a:b], far[a:b], S, t_rand)
            x_skel, mask, bw = ops.motion_field(scale_Rs[0], Ts[0], vol_j, bbox_min, bbox_scale,
                                                rays_o=o_c, rays_d=d_c, z=z, want_weights=heavy)
            n = b - a
            xs_flat = x_skel.reshape(-1, 3)
            keep_idx, count = None, None
            if cull_eps > 0.0 and torch.is_grad_enabled() and cull_train and precision == _lib.PRECISION_BF16:
                # the same skipping inside a training step, without telling the host how many samples
                # survive (fixed shapes: CUDA-graph capturable): the kept samples are compacted to the front
                # of a P-row buffer and the tcgen05 kernels read the count from device memory.  A skipped
                # sample contributes to no gradient; with mask < eps its true contributions are below
                # eps * O(1) (d loss / d sigma carries the factor mask, and d loss / d mask reaches the
                # parameters only through softmax / trilinear factors that are themselves below eps).
                P = xs_flat.shape[0]
                keep = mask.detach().reshape(-1) >= cull_eps
                pos = torch.cumsum(keep.to(torch.int64), 0) - 1
                count = (pos[-1:] + 1).contiguous()
                slot = torch.where(keep, pos, torch.full_like(pos, P))
                buf = torch.full((P + 1,), -1, dtype=torch.int64, device=dev)
                buf.scatter_(0, slot, torch.arange(P, dtype=torch.int64, device=dev))
                keep_idx = buf[:P].contiguous()
                self.last_cull_count = count
                xs_in = ops.gather_rows(xs_flat, keep_idx)
            elif cull_eps > 0.0 and not torch.is_grad_enabled():
                # empty-space skipping (opt-in): a sample's alpha is (1 - exp(-sigma delta)) *
                # mask <= mask, so dropping the samples with mask < eps changes no compositing weight by
                # more than eps and no output by more than 2 * S * eps.  They keep xyz = x_skel, offset = 0
                # and raw = 0 (sigma = 0).  torch.nonzero synchronises with the host: render path only.
                keep_idx = torch.nonzero(mask.reshape(-1) >= cull_eps)[:, 0]
                n_kept += int(keep_idx.numel())
                xs_in = ops.gather_rows(xs_flat, keep_idx)
            else:
                xs_in = xs_flat
            raw, xyz, offs = ops.canonical_field(self._spec, xs_in, cond, hann, nr_w, nr_b,
                                                 cn_w, cn_b, ignore_non_rigid=ignore_nr, precision=precision,
                                                 packed=packed, count=count)
            if keep_idx is not None:
                raw = ops.scatter_rows(raw, keep_idx, torch.zeros(n * S, 4, device=dev))
                xyz = ops.scatter_rows(xyz, keep_idx, xs_flat)
                offs = ops.scatter_rows(offs, keep_idx, torch.zeros(n * S, 3, device=dev))
            xyz = xyz.view(n, S, 3)
            rgb, acc, depth, weights, cnl_xyz, cnl_rgb, w_max, rgbs = ops.composite(
                raw.view(n, S, 4), mask, z, d_c, xyz, bgcolor, heavy=heavy)
            res = {'rgb': rgb, 'alpha': acc, 'depth': depth, 'weights_on_rays': weights, 'xyz_on_rays': xyz,
                   'cnl_xyz': cnl_xyz, 'cnl_rgb': cnl_rgb, 'cnl_weight': w_max, 'offsets': offs.view(n, S, 3)}
            if heavy:
                res['rgb_on_rays'] = rgbs
                res['backward_motion_weights'] = bw
            outs.append(res)

        if not outs:
            raise RuntimeError('humannerf_b200.Network.forward: no rays')
        if cull_eps > 0.0 and not torch.is_grad_enabled():
            self.last_cull_fraction = 1.0 - n_kept / float(R * S)
        all_ret = {k: (outs[0][k] if len(outs) == 1 else torch.cat([o[k] for o in outs], dim=0)) for k in outs[0]}
        for k in all_ret:
            all_ret[k] = all_ret[k].reshape(lead_shape + list(all_ret[k].shape[1:]))
        return all_ret
