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
                     n_patches: int, patch_size: int, rng: np.random.RandomState) -> Dict[str, np.ndarray]:
    """rays_o / rays_d / near / far are the COMPACTED per-ray arrays (ray_mask order)."""
    sel, masks, div, xy = sample_patches(ray_mask, subject_mask, H, W, n_patches, patch_size, rng)
    targets = np.stack([image[y:y + patch_size, x:x + patch_size] for x, y in xy], axis=0)
    return {
        'rays': np.stack([rays_o[sel], rays_d[sel], rays_d[sel]], axis=0),
        'near': near[sel].reshape(-1, 1).astype(np.float32),
        'far': far[sel].reshape(-1, 1).astype(np.float32),
        'patch_masks': masks,
        'patch_div_indices': div,
        'target_patches': targets.astype(np.float32),
    }
