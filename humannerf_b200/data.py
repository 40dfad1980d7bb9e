"""Host-side batch assembly for a training step, mirroring what the reference Dataset hands to the
network (`core/data/human_nerf/train.py:235-340, 450-476, 592-650`): pick N patches (80 % centred on
the subject, 20 % on bbox background), intersect each window with the ray mask, and gather the
selected rays / target pixels.  numpy on the host, like the reference (DataLoader workers)."""
from __future__ import annotations

from typing import Dict

import numpy as np


def sample_patches(ray_mask: np.ndarray, subject_mask: np.ndarray, H: int, W: int, n_patches: int,
                   patch_size: int, rng: np.random.RandomState, sample_subject_ratio: float = 0.8):
    """Returns select_inds (indices into the compacted ray list), patch_masks (N,P,P) bool,
    div_indices (N+1,), xy_min (N,2).  Two host randoms per patch, drawn in the reference's order
    (`rand()` then `choice(n)`), so a seeded RandomState reproduces the reference's picks."""
    ray_mask = np.asarray(ray_mask, dtype=bool).reshape(-1)
    bbox_mask = ray_mask.reshape(H, W)
    subject_mask = np.asarray(subject_mask, dtype=bool)
    background = bbox_mask & ~subject_mask
    compact = np.cumsum(ray_mask) - 1
    sel, masks, div, xy = [], [], [0], []
    for _ in range(n_patches):
        cand = subject_mask if rng.rand(1)[0] < sample_subject_ratio else background
        ys, xs = np.where(cand)
        k = rng.choice(ys.shape[0], size=[1], replace=False)[0]
        x0 = int(np.clip(xs[k] - patch_size // 2, 0, W - patch_size))
        y0 = int(np.clip(ys[k] - patch_size // 2, 0, H - patch_size))
        window = np.zeros((H, W), dtype=bool)
        window[y0:y0 + patch_size, x0:x0 + patch_size] = True
        hit = window.reshape(-1) & ray_mask
        sel.append(compact[np.where(hit)[0]])
        masks.append(hit.reshape(H, W)[y0:y0 + patch_size, x0:x0 + patch_size])
        div.append(div[-1] + sel[-1].shape[0])
        xy.append((x0, y0))
    return np.concatenate(sel), np.stack(masks), np.asarray(div), np.asarray(xy)


def make_train_batch(rays_o: np.ndarray, rays_d: np.ndarray, near: np.ndarray, far: np.ndarray,
                     ray_mask: np.ndarray, subject_mask: np.ndarray, image: np.ndarray, H: int, W: int,
                     n_patches: int, patch_size: int, rng: np.random.RandomState,
                     pad_rays: bool = False) -> Dict[str, np.ndarray]:
    """rays_o / rays_d / near / far are the COMPACTED per-ray arrays (ray_mask order).

    `scatter_index[r]` is the flat pixel (patch, y, x) ray r lands on (what `patch_imgs[i, patch_masks[i]] =
    rgbs[div[i]:div[i+1]]` does in trainer.py:35), computed here on the host so the step needs no
    device-side `nonzero`.  With `pad_rays` every per-ray array is padded to n_patches * patch_size**2
    rows (zero rays, near = far = 0, scatter_index = -1: the scatter skips them, so they get no
    gradient) - fixed shapes for the CUDA-graph train step; `n_rays` is the number of real rays."""
    sel, masks, div, xy = sample_patches(ray_mask, subject_mask, H, W, n_patches, patch_size, rng)
    targets = np.stack([image[y:y + patch_size, x:x + patch_size] for x, y in xy], axis=0)
    rays = np.stack([rays_o[sel], rays_d[sel], rays_d[sel]], axis=0)
    near_s = near[sel].reshape(-1, 1).astype(np.float32)
    far_s = far[sel].reshape(-1, 1).astype(np.float32)
    index = np.nonzero(masks.reshape(-1))[0].astype(np.int64)
    n = int(sel.shape[0])
    if pad_rays:
        cap = n_patches * patch_size * patch_size
        rays = np.concatenate([rays, np.zeros((3, cap - n, 3), rays.dtype)], axis=1)
        near_s = np.concatenate([near_s, np.zeros((cap - n, 1), np.float32)], axis=0)
        far_s = np.concatenate([far_s, np.zeros((cap - n, 1), np.float32)], axis=0)
        index = np.concatenate([index, np.full(cap - n, -1, np.int64)])
    return {
        'rays': rays,
        'near': near_s,
        'far': far_s,
        'patch_masks': masks,
        'patch_div_indices': div,
        'target_patches': targets.astype(np.float32),
        'scatter_index': index,
        'n_rays': np.asarray(n, dtype=np.int64),
    }


class DevicePatchSampler:
    """Patch selection on the device (a3 / SURVEY 8f row 2): the frame's compacted rays, masks and image stay
    resident in HBM; per step the host only draws the reference's two randoms per patch (`draw`) and ONE kernel
    (`hn_select_patches`, csrc/patches.cu) writes the whole padded batch.  Integer outputs (`select_inds`,
    `patch_masks`, `div_indices`, `xy_min`) are bit-exact with `sample_patches` / the reference for the same draws.

    ray_mask (H*W) uint8 / bool device tensor (= the reference's `bbox_mask` flattened), subject_mask (H, W) device
    tensor, rays_o / rays_d (R, 3), near / far (R) or (R, 1): the COMPACTED fp32 ray arrays (`ops.compact_rays`),
    image (H, W, 3) fp32 on the device."""

    def __init__(self, ray_mask, subject_mask, rays_o, rays_d, near, far, image, H: int, W: int, n_patches: int,
                 patch_size: int, sample_subject_ratio: float = 0.8):
        import ctypes as C
        import torch
        from humannerf_b200 import _lib
        self._lib, self._torch = _lib, torch
        dev = rays_o.device
        if dev.type != 'cuda':
            raise RuntimeError('DevicePatchSampler: tensors must live on a CUDA device (no CPU fallback)')
        self.H, self.W, self.n_patches, self.patch, self.ratio = H, W, int(n_patches), int(patch_size), float(sample_subject_ratio)
        u8 = lambda t: t.reshape(-1).to(torch.uint8).contiguous()
        self.ray_mask, self.subject = u8(ray_mask), u8(subject_mask)
        f32 = lambda t: t.to(torch.float32).contiguous()
        self.rays_o, self.rays_d = f32(rays_o), f32(rays_d)
        self.near, self.far = f32(near).reshape(-1), f32(far).reshape(-1)
        self.image = f32(image).reshape(H, W, 3)
        n = H * W
        self.cum = torch.empty(3, n, dtype=torch.int32, device=dev)      # ray / subject / background prefix counts
        counts = torch.zeros(3, dtype=torch.int32, device=dev)
        lib = _lib.load()
        with torch.cuda.device(dev):
            _lib.check(lib.hn_patch_prepare(H, W, _lib.ptr(self.ray_mask), _lib.ptr(self.subject), _lib.ptr(self.cum[0]),
                                            _lib.ptr(self.cum[1]), _lib.ptr(self.cum[2]), _lib.ptr(counts),
                                            _lib.stream_ptr()), 'patch_prepare')
        # once per FRAME: the host needs the candidate counts to draw np.random.choice(n) like the reference
        self.n_rays_frame, self.n_subject, self.n_background = [int(v) for v in counts.tolist()]
        if self.n_rays_frame != self.rays_o.shape[0]:
            raise RuntimeError(f'DevicePatchSampler: ray_mask selects {self.n_rays_frame} rays, got {self.rays_o.shape[0]}')
        self.cap = self.n_patches * self.patch * self.patch

    def draw(self, rng) -> 'np.ndarray':
        """The reference's host randoms for one step, in its order (train.py:263, :307): per patch `rand(1)` then
        `choice(n_candidates)`.  Returns a (2, n_patches) float64 array [u; k] (k is exact in float64)."""
        out = np.empty((2, self.n_patches), dtype=np.float64)
        for i in range(self.n_patches):
            u = rng.rand(1)[0]
            n = self.n_subject if u < self.ratio else self.n_background
            out[0, i] = u
            out[1, i] = rng.choice(n, size=[1], replace=False)[0]
        return out

    def alloc(self):
        """Output buffers of one step (reusable; `select` overwrites every element)."""
        torch = self._torch
        dev, n, P, cap = self.rays_o.device, self.n_patches, self.patch, self.cap
        return {'rays': torch.empty(3, cap, 3, device=dev), 'near': torch.empty(cap, 1, device=dev),
                'far': torch.empty(cap, 1, device=dev), 'scatter_index': torch.empty(cap, dtype=torch.int64, device=dev),
                'target_patches': torch.empty(n, P, P, 3, device=dev),
                'select_inds': torch.empty(cap, dtype=torch.int32, device=dev),
                'patch_masks': torch.empty(n, P, P, dtype=torch.uint8, device=dev),
                'patch_div_indices': torch.empty(n + 1, dtype=torch.int32, device=dev),
                'xy_min': torch.empty(n, 2, dtype=torch.int32, device=dev),
                'n_rays': torch.empty(1, dtype=torch.int64, device=dev)}

    def select(self, u, k, out=None):
        """u (n_patches,) float64 and k (n_patches,) int64 DEVICE tensors -> the padded batch (device tensors).
        Asynchronous, no host round trip: capturable in a CUDA graph."""
        torch, _lib = self._torch, self._lib
        out = self.alloc() if out is None else out
        lib = _lib.load()
        ptr = _lib.ptr
        assert u.dtype == torch.float64 and k.dtype == torch.int64 and u.numel() == self.n_patches == k.numel()
        with torch.cuda.device(self.rays_o.device):
            _lib.check(lib.hn_select_patches(
                self.H, self.W, self.n_patches, self.patch, self.ratio, ptr(self.ray_mask), ptr(self.cum[0]),
                ptr(self.cum[1]), ptr(self.cum[2]), ptr(u), ptr(k), ptr(self.rays_o), ptr(self.rays_d), ptr(self.near),
                ptr(self.far), ptr(self.image), ptr(out['select_inds']), ptr(out['patch_masks']),
                ptr(out['patch_div_indices']), ptr(out['xy_min']), ptr(out['target_patches']), ptr(out['rays']),
                ptr(out['near']), ptr(out['far']), ptr(out['scatter_index']), ptr(out['n_rays']), _lib.stream_ptr()),
                'select_patches')
        return out
