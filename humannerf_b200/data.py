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
