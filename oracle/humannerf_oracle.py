"""CPU oracle for HumanNeRF's per-ray rendering hot path.  TEST INFRASTRUCTURE.

This file restates, in plain torch-CPU fp32 (and numpy float64 for the ray
set-up, as the reference does), the algorithm of the reference's hot path so
that the CUDA kernels can be checked against it.  It is *only* the checker:
nothing under `humannerf_b200/` imports it; only `tests/`,
`__graft_entry__.smoke()` and `bench.py`'s cpu-baseline / `--impl reference`
legs do.

Parity pin: the reference has no tests or golden vectors of its own (SURVEY.md
section 4), so this oracle is pinned against outputs of the reference itself,
produced in the build container by `tests/golden/make_golden.py` (which imports
the unmodified reference from /root/reference) and committed under
`tests/golden/*.npz`; `tests/test_oracle_golden.py` replays them.

Each function cites the reference code it follows (paths relative to the
reference root).  Everything is functional: parameters come in as a
`state_dict`-style mapping with the reference's parameter names.
"""
from __future__ import annotations

import math
from typing import Dict, Mapping, Optional, Sequence, Tuple

import numpy as np
import torch
import torch.nn.functional as F

SMPL_PARENT = (-1, 0, 0, 0, 1, 2, 3, 4, 5, 6, 7, 8, 9, 9, 9, 12, 13, 14, 16, 17, 18, 19, 20, 21)


# ---------------------------------------------------------------------------
# a1/a2: rays from the camera and ray / bbox intersection  (numpy, float64)
# ---------------------------------------------------------------------------
def rays_from_camera(H: int, W: int, K: np.ndarray, R: np.ndarray, T: np.ndarray):
    """`core/utils/camera_util.py:132-159` get_rays_from_KRT.
    pixel (x, y, 1) -> camera via K^-1 -> world via R^T (x_cam - T); the origin
    is -R^T T; directions are *not* normalised."""
    origin = -(R.T @ T.reshape(3)).ravel()
    xs, ys = np.meshgrid(np.arange(W, dtype=np.float32), np.arange(H, dtype=np.float32),
                         indexing='xy')
    pix = np.stack([xs, ys, np.ones_like(xs)], axis=2)
    cam = pix @ np.linalg.inv(K).T
    world = (cam - T.reshape(3)) @ R
    rays_d = world - origin[None, None]
    rays_o = np.broadcast_to(origin, rays_d.shape)
    return rays_o, rays_d


def rays_bbox_near_far(bounds: np.ndarray, ray_o: np.ndarray, ray_d: np.ndarray):
    """`core/utils/camera_util.py:162-208` rays_intersect_3d_bbox.

    bounds (2,3) = [min_xyz, max_xyz], padded by 0.01.  Direction components
    with |d| < 1e-5 are replaced by 1e-5 (the reference does this in place, so
    the returned `ray_d` is what the network later sees).  A ray is kept when
    exactly two of its six plane hits lie inside the (eps-grown) box.
    Returns near, far (over kept rays), mask (over all rays), patched ray_d."""
    bounds = np.asarray(bounds) + np.array([-0.01, 0.01])[:, None]
    ray_d = np.array(ray_d, copy=True)
    ray_d[np.abs(ray_d) < 1e-5] = 1e-5
    t_hit = ((bounds[None] - ray_o[:, None]) / ray_d[:, None]).reshape(-1, 6)
    p_hit = t_hit[..., None] * ray_d[:, None] + ray_o[:, None]          # (N, 6, 3)
    lo, hi = bounds[0] - 1e-6, bounds[1] + 1e-6
    inside = np.ones(p_hit.shape[:2], dtype=bool)
    for a in range(3):
        inside &= (p_hit[..., a] >= lo[a]) & (p_hit[..., a] <= hi[a])
    mask = inside.sum(-1) == 2
    pts = p_hit[mask][inside[mask]].reshape(-1, 2, 3)
    o, d = ray_o[mask], ray_d[mask]
    dn = np.linalg.norm(d, axis=1)
    d0 = np.linalg.norm(pts[:, 0] - o, axis=1) / dn
    d1 = np.linalg.norm(pts[:, 1] - o, axis=1) / dn
    return np.minimum(d0, d1), np.maximum(d0, d1), mask, ray_d


# ---------------------------------------------------------------------------
# a3: patch ray selection (integer / bool index math)
# ---------------------------------------------------------------------------
def patch_ray_indices(ray_mask: np.ndarray, subject_mask: np.ndarray, H: int, W: int,
                      patch_size: int, u_subject: Sequence[float], u_center: Sequence[float],
                      sample_subject_ratio: float = 0.8):
    """`core/data/human_nerf/train.py:235-340`.

    The reference draws two host randoms per patch: `np.random.rand()` (subject
    vs. bbox-background) and `np.random.choice(n_candidates)`.  Here both are
    inputs: `u_subject[i]` in [0,1) and `u_center[i]` in [0,1) mapped to
    candidate index floor(u * n) - tests that replay the reference pass the
    index it drew via `u_center = (idx + 0.5) / n`.
    Returns select_inds, patch_masks (N,P,P) bool, div_indices (N+1,), xy_min (N,2)."""
    ray_mask = np.asarray(ray_mask, dtype=bool).reshape(-1)
    bbox_mask = ray_mask.reshape(H, W)
    subject_mask = np.asarray(subject_mask, dtype=bool)
    background = bbox_mask & ~subject_mask
    compact = np.cumsum(ray_mask) - 1
    sel, masks, div, xy = [], [], [0], []
    for us, uc in zip(u_subject, u_center):
        cand = subject_mask if us < sample_subject_ratio else background
        ys, xs = np.where(cand)
        k = min(int(uc * ys.shape[0]), ys.shape[0] - 1)
        x0 = int(np.clip(xs[k] - patch_size // 2, 0, W - patch_size))
        y0 = int(np.clip(ys[k] - patch_size // 2, 0, H - patch_size))
        window = np.zeros((H, W), dtype=bool)
        window[y0:y0 + patch_size, x0:x0 + patch_size] = True
        hit = window.reshape(-1) & ray_mask
        sel.append(compact[np.where(hit)[0]])
        masks.append(hit.reshape(H, W)[y0:y0 + patch_size, x0:x0 + patch_size])
        div.append(div[-1] + sel[-1].shape[0])
        xy.append((x0, y0))
    return (np.concatenate(sel), np.stack(masks), np.asarray(div), np.asarray(xy))


def unpack_patches(rgb: torch.Tensor, patch_masks: torch.Tensor, bgcolor01: torch.Tensor,
                   div_indices) -> torch.Tensor:
    """`core/train/trainers/human_nerf/trainer.py:28-37` _unpack_imgs: scatter
    rays into (N,P,P,3) patches pre-filled with the background colour."""
    n, p = patch_masks.shape[0], patch_masks.shape[1]
    out = bgcolor01.to(rgb).expand(n, p, p, 3).clone()
    for i in range(n):
        out[i, patch_masks[i]] = rgb[int(div_indices[i]):int(div_indices[i + 1])]
    return out


# ---------------------------------------------------------------------------
# a9: samples along the ray
# ---------------------------------------------------------------------------
def z_samples(near: torch.Tensor, far: torch.Tensor, n_samples: int,
              t_rand: Optional[torch.Tensor] = None) -> torch.Tensor:
    """`network.py:455-471`.  near/far (R,1).  z = near*(1-t) + far*t with t =
    linspace(0,1,S); optional stratified jitter between neighbouring mid-points
    with the uniform noise `t_rand` (R,S) supplied by the caller."""
    t = torch.linspace(0.0, 1.0, steps=n_samples).to(near)
    z = (near * (1.0 - t) + far * t).expand(near.shape[0], n_samples)
    if t_rand is not None:
        mid = 0.5 * (z[:, 1:] + z[:, :-1])
        hi = torch.cat([mid, z[:, -1:]], dim=-1)
        lo = torch.cat([z[:, :1], mid], dim=-1)
        z = lo + (hi - lo) * t_rand
    return z


# ---------------------------------------------------------------------------
# a5/a6: pose refinement and motion bases (per frame)
# ---------------------------------------------------------------------------
def rodrigues(rvec: torch.Tensor) -> torch.Tensor:
    """`core/utils/network_util.py:57-83`: theta = sqrt(1e-5 + |r|^2) (so a zero
    vector does *not* give the identity exactly)."""
    theta = torch.sqrt(1e-5 + (rvec * rvec).sum(dim=1))
    u = rvec / theta[:, None]
    c, s = torch.cos(theta), torch.sin(theta)
    x, y, z = u[:, 0], u[:, 1], u[:, 2]
    rows = [x * x + (1 - x * x) * c, x * y * (1 - c) - z * s, x * z * (1 - c) + y * s,
            x * y * (1 - c) + z * s, y * y + (1 - y * y) * c, y * z * (1 - c) - x * s,
            x * z * (1 - c) - y * s, y * z * (1 - c) + x * s, z * z + (1 - z * z) * c]
    return torch.stack(rows, dim=1).view(-1, 3, 3)


def mlp_chain(x: torch.Tensor, sd: Mapping[str, torch.Tensor], prefix: str,
              layer_ids: Sequence[int], relu_last: bool) -> torch.Tensor:
    for n, i in enumerate(layer_ids):
        x = F.linear(x, sd[f'{prefix}.{i}.weight'], sd[f'{prefix}.{i}.bias'])
        if relu_last or n + 1 < len(layer_ids):
            x = torch.relu(x)
    return x


def refine_pose(sd, dst_Rs: torch.Tensor, dst_posevec: torch.Tensor) -> torch.Tensor:
    """`pose_decoders/mlp_delta_body_pose.py:35-41` + `network.py:612-616,667-688`:
    69-d pose -> MLP -> 23 axis-angle corrections; R_i <- R_i @ dR_i for i >= 1."""
    rvec = mlp_chain(dst_posevec[None], sd, 'pose_decoder.block_mlps', (0, 2, 4, 6, 8), False)
    dR = rodrigues(rvec.view(-1, 3))
    return torch.cat([dst_Rs[:1], torch.matmul(dst_Rs[1:], dR)], dim=0)


def motion_bases(dst_Rs: torch.Tensor, dst_Ts: torch.Tensor, cnl_gtfms: torch.Tensor):
    """`core/utils/network_util.py:97-156` MotionBasisComputer: forward
    kinematics down the SMPL tree, then canonical_global @ inverse(posed_global)
    per joint.  Returns scale_Rs (24,3,3), Ts (24,3)."""
    n = dst_Rs.shape[0]
    local = torch.zeros(n, 4, 4, dtype=dst_Rs.dtype, device=dst_Rs.device)
    local[:, :3, :3] = dst_Rs
    local[:, :3, 3] = dst_Ts
    local[:, 3, 3] = 1.0
    glob = [local[0]]
    for i in range(1, n):
        glob.append(glob[SMPL_PARENT[i]] @ local[i])
    f = cnl_gtfms @ torch.inverse(torch.stack(glob))
    return f[:, :3, :3], f[:, :3, 3]


def decode_motion_weight_volume(sd, prior: torch.Tensor) -> torch.Tensor:
    """`mweight_vol_decoders/deconv_vol_decoder.py:25-33` + `network_util.py:12-50`:
    const embedding -> Linear+LeakyReLU(0.2) -> 5 x ConvTranspose3d(k4,s2,p1)
    (LeakyReLU between) -> softmax(logits + log prior) over the 25 channels."""
    p = 'mweight_vol_decoder'
    h = F.leaky_relu(F.linear(sd[f'{p}.const_embedding'][None],
                              sd[f'{p}.decoder.block_mlp.0.weight'],
                              sd[f'{p}.decoder.block_mlp.0.bias']), 0.2)
    h = h.view(1, -1, 1, 1, 1)
    convs = sorted({int(k.split('.')[3]) for k in sd
                    if k.startswith(f'{p}.decoder.block_conv.') and k.endswith('.weight')})
    for n, i in enumerate(convs):
        h = F.conv_transpose3d(h, sd[f'{p}.decoder.block_conv.{i}.weight'],
                               sd[f'{p}.decoder.block_conv.{i}.bias'], stride=2, padding=1)
        if n + 1 < len(convs):
            h = F.leaky_relu(h, 0.2)
    return torch.softmax(h + torch.log(prior[None]), dim=1)[0]


# ---------------------------------------------------------------------------
# a10: motion field (trilinear lookup of 24 blend-weight channels + inverse LBS)
# ---------------------------------------------------------------------------
def trilinear_zero_pad(vol: torch.Tensor, g: torch.Tensor) -> torch.Tensor:
    """Trilinear lookup, `align_corners=True`, zero padding, exactly what the
    reference asks of `F.grid_sample` (`network.py:411-413`).
    vol (J,D,H,W) indexed [z][y][x]; g (J,P,3) in [-1,1] as (x,y,z). -> (J,P)."""
    J, D, Hh, Ww = vol.shape
    fx = (g[..., 0] + 1.0) * 0.5 * (Ww - 1)
    fy = (g[..., 1] + 1.0) * 0.5 * (Hh - 1)
    fz = (g[..., 2] + 1.0) * 0.5 * (D - 1)
    x0, y0, z0 = torch.floor(fx), torch.floor(fy), torch.floor(fz)
    tx, ty, tz = fx - x0, fy - y0, fz - z0
    flat = vol.reshape(J, -1)
    out = torch.zeros_like(fx)
    for dz in (0, 1):
        for dy in (0, 1):
            for dx in (0, 1):
                xi, yi, zi = x0 + dx, y0 + dy, z0 + dz
                w = ((tx if dx else 1 - tx) * (ty if dy else 1 - ty) * (tz if dz else 1 - tz))
                ok = (xi >= 0) & (xi < Ww) & (yi >= 0) & (yi < Hh) & (zi >= 0) & (zi < D)
                idx = (zi.clamp(0, D - 1) * Hh + yi.clamp(0, Hh - 1)) * Ww + xi.clamp(0, Ww - 1)
                val = torch.gather(flat, 1, idx.long())
                out = out + torch.where(ok, val * w, torch.zeros_like(val))
    return out


def sample_motion_fields(pts: torch.Tensor, scale_Rs: torch.Tensor, Ts: torch.Tensor,
                         vol: torch.Tensor, bbox_min: torch.Tensor, bbox_scale: torch.Tensor):
    """`network.py:392-444`.  pts (P,3) -> x_skel (P,3), mask (P,1) = sum of the
    24 sampled weights (background channel dropped), weights (P,24)."""
    J = vol.shape[0] - 1
    pos = torch.einsum('jab,pb->jpa', scale_Rs[:J], pts) + Ts[:J, None, :]
    g = (pos - bbox_min) * bbox_scale - 1.0
    w = trilinear_zero_pad(vol[:J], g)                                   # (J,P)
    wsum = w.sum(dim=0)[:, None]
    x_skel = (w[..., None] * pos).sum(dim=0) / wsum.clamp(min=1e-4)
    return x_skel, wsum, w.t().contiguous()


# ---------------------------------------------------------------------------
# a11: positional encodings
# ---------------------------------------------------------------------------
def fourier_pe(x: torch.Tensor, n_freqs: int, weights: Optional[torch.Tensor] = None,
               include_input: bool = True) -> torch.Tensor:
    """`embedders/fourier.py:9-38`: [x, sin(2^k x), cos(2^k x)]_k; with `weights`
    (and include_input=False) it is the Hann-windowed variant of
    `embedders/hannw_fourier.py:38-49`."""
    parts = [x] if include_input else []
    freqs = 2.0 ** torch.linspace(0.0, n_freqs - 1, steps=n_freqs)
    for k in range(n_freqs):
        s, c = torch.sin(x * freqs[k]), torch.cos(x * freqs[k])
        if weights is not None:
            s, c = weights[k] * s, weights[k] * c
        parts += [s, c]
    return torch.cat(parts, dim=-1)


def hann_weights(iter_val, n_freqs: int, kick_in_iter: float, full_band_iter: float) -> torch.Tensor:
    """`embedders/hannw_fourier.py:24-40`: alpha = m * max(iter - T_s, 0) / (T_e - T_s);
    w_k = (1 - cos(pi * clamp(alpha - k, 0, 1))) / 2."""
    it = torch.as_tensor(iter_val, dtype=torch.float32).reshape(-1)[0]
    t = torch.clamp(it - float(kick_in_iter), min=0.0)
    span = float(full_band_iter) - float(kick_in_iter)
    alpha = torch.tensor(float(n_freqs)) if span == 0 else n_freqs * t / span
    k = torch.arange(n_freqs, dtype=torch.float32, device=it.device)
    return (1.0 - torch.cos(math.pi * torch.clamp(alpha - k, 0.0, 1.0))) / 2.0


# ---------------------------------------------------------------------------
# a12/a13: the two MLPs
# ---------------------------------------------------------------------------
def non_rigid_mlp(sd, pe: torch.Tensor, pos: torch.Tensor, cond: torch.Tensor,
                  skip_at: int = 8, prefix: str = 'non_rigid_mlp.module.block_mlps'):
    """`non_rigid_motion_mlps/mlp_offset.py:74-114`: h = [cond | pe]; Linear+ReLU
    blocks with the pe re-appended ([h | pe]) before module index `skip_at`;
    final Linear(.,3) without activation; xyz = pos + offset."""
    ids = sorted({int(k[len(prefix) + 1:].split('.')[0]) for k in sd if k.startswith(prefix + '.')})
    h = torch.cat([cond.expand(pe.shape[0], -1), pe], dim=-1)
    for i in ids:
        if i == skip_at:
            h = torch.cat([h, pe], dim=-1)
        h = F.linear(h, sd[f'{prefix}.{i}.weight'], sd[f'{prefix}.{i}.bias'])
        if i != ids[-1]:
            h = torch.relu(h)
    return pos + h, h


def canonical_mlp(sd, pe: torch.Tensor, skip_at: int = 10, prefix: str = 'cnl_mlp.module'):
    """`canonical_mlps/mlp_rgb_sigma.py:132-206` (default branch): Linear+ReLU
    trunk with [pe | h] before module index `skip_at`, then Linear(.,4) -> raw
    (r,g,b,sigma) without activation."""
    trunk = f'{prefix}.pts_linears'
    ids = sorted({int(k[len(trunk) + 1:].split('.')[0]) for k in sd if k.startswith(trunk + '.')})
    h = pe
    for i in ids:
        if i == skip_at:
            h = torch.cat([pe, h], dim=-1)
        h = torch.relu(F.linear(h, sd[f'{trunk}.{i}.weight'], sd[f'{trunk}.{i}.bias']))
    return F.linear(h, sd[f'{prefix}.output_linear.0.weight'], sd[f'{prefix}.output_linear.0.bias'])


# ---------------------------------------------------------------------------
# a15: alpha compositing
# ---------------------------------------------------------------------------
def composite(raw: torch.Tensor, mask: torch.Tensor, z: torch.Tensor, rays_d: torch.Tensor,
              xyz: torch.Tensor, bgcolor255: torch.Tensor) -> Dict[str, torch.Tensor]:
    """`network.py:355-388` _raw2outputs.  raw (R,S,4), mask (R,S,1), z (R,S)."""
    R = raw.shape[0]
    delta = torch.cat([z[:, 1:] - z[:, :-1], torch.full((R, 1), 1e10).to(z)], dim=-1)
    delta = delta * torch.norm(rays_d[:, None, :], dim=-1)
    rgb = torch.sigmoid(raw[..., :3])
    alpha = (1.0 - torch.exp(-torch.relu(raw[..., 3]) * delta)) * mask[..., 0]
    trans = torch.cumprod(torch.cat([torch.ones(R, 1).to(alpha), 1.0 - alpha + 1e-10], dim=-1),
                          dim=-1)[:, :-1]
    w = alpha * trans
    acc = w.sum(-1)
    rgb_map = (w[..., None] * rgb).sum(-2) + (1.0 - acc[:, None]) * bgcolor255[None, :] / 255.0
    depth = (w * z).sum(-1)
    w_max, idx = w.max(dim=1)
    take = idx[:, None, None].expand(-1, 1, 3)
    return {'rgb': rgb_map, 'alpha': acc, 'weights_on_rays': w, 'depth': depth,
            'cnl_xyz': torch.gather(xyz, 1, take)[:, 0], 'cnl_rgb': torch.gather(rgb, 1, take)[:, 0],
            'cnl_weight': w_max, 'rgb_on_rays': rgb}


# ---------------------------------------------------------------------------
# a4: the whole forward
# ---------------------------------------------------------------------------
def network_forward(sd: Mapping[str, torch.Tensor], rays: torch.Tensor, near: torch.Tensor,
                    far: torch.Tensor, frame: Mapping[str, torch.Tensor], iter_val,
                    n_samples: int = 128, t_rand: Optional[torch.Tensor] = None,
                    nr_multires: int = 6, cnl_multires: int = 10,
                    nr_kick_in: float = 10000, nr_full_band: float = 50000,
                    pose_kick_in: float = 0, pose_decoder_off: bool = False,
                    ignore_non_rigid: bool = False, chunk: int = 32768,
                    return_intermediates: bool = False) -> Dict[str, torch.Tensor]:
    """`network.py:647-789` Network.forward for the default component set
    (global condition code, axis-angle posevec, no view dirs, single head).

    rays (3,R,3) = [o, d, d_camera]; near/far (R,1); `frame` holds dst_Rs,
    dst_Ts, cnl_gtfms, motion_weights_priors, dst_posevec, cnl_bbox_min_xyz,
    cnl_bbox_scale_xyz, bgcolor (0-255)."""
    it = float(torch.as_tensor(iter_val).reshape(-1)[0])
    dst_Rs = frame['dst_Rs']
    if it >= pose_kick_in and not pose_decoder_off:
        dst_Rs = refine_pose(sd, dst_Rs, frame['dst_posevec'])
    cond = frame['dst_posevec'][None]
    if it < nr_kick_in:
        cond = torch.zeros_like(cond)
    hann = hann_weights(iter_val, nr_multires, nr_kick_in, nr_full_band)
    scale_Rs, Ts = motion_bases(dst_Rs, frame['dst_Ts'], frame['cnl_gtfms'])
    vol = decode_motion_weight_volume(sd, frame['motion_weights_priors'])

    rays_o, rays_d = rays[0].reshape(-1, 3).float(), rays[1].reshape(-1, 3).float()
    outs = []
    for a in range(0, rays_o.shape[0], chunk):
        o, d = rays_o[a:a + chunk], rays_d[a:a + chunk]
        z = z_samples(near[a:a + chunk], far[a:a + chunk], n_samples,
                      None if t_rand is None else t_rand[a:a + chunk])
        pts = o[:, None, :] + d[:, None, :] * z[..., None]
        x_skel, mask, bw = sample_motion_fields(pts.reshape(-1, 3), scale_Rs, Ts, vol,
                                                frame['cnl_bbox_min_xyz'],
                                                frame['cnl_bbox_scale_xyz'])
        if ignore_non_rigid:
            xyz, offs = x_skel, torch.zeros_like(x_skel)
        else:
            pe_nr = fourier_pe(x_skel, nr_multires, weights=hann, include_input=False)
            xyz, offs = non_rigid_mlp(sd, pe_nr, x_skel, cond)
        raw = canonical_mlp(sd, fourier_pe(xyz, cnl_multires))
        r, s = z.shape
        res = composite(raw.view(r, s, 4), mask.view(r, s, 1), z, d, xyz.view(r, s, 3),
                        frame['bgcolor'])
        res.update({'xyz_on_rays': xyz.view(r, s, 3), 'offsets': offs.view(r, s, 3),
                    'backward_motion_weights': bw.view(r, s, -1)})
        if return_intermediates:
            res.update({'z_vals': z, 'x_skel': x_skel.view(r, s, 3), 'mask': mask.view(r, s, 1),
                        'raw': raw.view(r, s, 4)})
        outs.append(res)
    out = {k: torch.cat([o_[k] for o_ in outs], dim=0) for k in outs[0]}
    if return_intermediates:
        out.update({'motion_scale_Rs': scale_Rs, 'motion_Ts': Ts, 'motion_weights_vol': vol})
    return out


def train_loss_mse(rgb: torch.Tensor, patch_masks, bgcolor01, targets, div_indices) -> torch.Tensor:
    """MSE term of `trainer.py:97-113,115-175` (the LPIPS term needs pretrained
    VGG16 weights that are not available offline - SURVEY.md section 8c)."""
    return torch.mean((unpack_patches(rgb, patch_masks, bgcolor01, div_indices) - targets) ** 2)
