"""Frame rendering through the device path: what `run.py` does per frame with numpy rays
(`core/data/human_nerf/{tpose,freeview,movement}.py` -> `get_rays_from_KRT` -> `rays_intersect_3d_bbox` ->
`rays[ray_mask]`, then `Network.forward`, then `unpack_to_image` run.py:49-65) with every stage on the GPU:

    hn_gen_rays -> hn_rays_bbox -> hn_compact_rays -> Network.forward (chunked) -> hn_scatter_rows into H x W

Nothing but the 4x4 camera and the per-frame pose tensors crosses the host boundary."""
from __future__ import annotations

from typing import Dict, Optional, Tuple

import numpy as np
import torch

from humannerf_b200 import ops
from humannerf_b200.config import get_cfg

FRAME_KEYS = ('dst_Rs', 'dst_Ts', 'cnl_gtfms', 'motion_weights_priors', 'dst_posevec', 'cnl_bbox_min_xyz',
              'cnl_bbox_max_xyz', 'cnl_bbox_scale_xyz', 'bgcolor')


def frame_rays(H: int, W: int, K: np.ndarray, E: np.ndarray, bounds: np.ndarray, device):
    """Compacted fp32 rays of the pixels that hit the (posed) bounding box, in pixel order:
    rays_o (R,3), rays_d (R,3), near (R,1), far (R,1), pixel_index (R) int32, ray_mask (H*W) uint8."""
    o64, d64 = ops.gen_rays(H, W, K, E, device)
    near, far, mask = ops.rays_bbox(bounds, o64, d64)
    o, d, nr, fa, pix = ops.compact_rays(mask, o64, d64, near, far)
    return o, d, nr, fa, pix, mask


def render_view(net, frame: Dict[str, torch.Tensor], K: np.ndarray, E: np.ndarray, H: int, W: int,
                bounds: np.ndarray, rank: int = 0, world: int = 1, gather: bool = True,
                iter_val=None) -> Tuple[torch.Tensor, torch.Tensor, int]:
    """One view: image (H*W, 3) with the background colour where no ray was cast, alpha (H*W), number of rays.
    With world > 1 the rays are split contiguously over the ranks and (gather=True) the per-ray results
    (rgb, alpha, depth: 20 B / ray) are all-gathered so that every rank holds the image."""
    from humannerf_b200.parallel import all_gather_rows, shard_bounds
    cfg = get_cfg()
    dev = frame['dst_Rs'].device
    o, d, nr, fa, pix, _ = frame_rays(H, W, K, E, bounds, dev)
    n = o.shape[0]
    a, b = shard_bounds(n, rank, world) if world > 1 else (0, n)
    kw = {k: frame[k] for k in FRAME_KEYS}
    out = net(rays=torch.stack([o[a:b], d[a:b], d[a:b]], 0), near=nr[a:b], far=fa[a:b],
              iter_val=cfg.eval_iter if iter_val is None else iter_val, **kw)
    packed = torch.cat([out['rgb'], out['alpha'][:, None], out['depth'][:, None]], dim=1)
    if world > 1 and gather:
        packed = all_gather_rows(packed, n)
        pix_local = pix
    else:
        pix_local = pix[a:b]
    bg = (frame['bgcolor'].float() / 255.0).reshape(1, 3)
    image = ops.scatter_rows(packed[:, :3], pix_local.long(), bg.expand(H * W, 3).contiguous())
    alpha = ops.scatter_rows(packed[:, 3:4], pix_local.long(), torch.zeros(H * W, 1, device=dev))
    return image, alpha[:, 0], n
