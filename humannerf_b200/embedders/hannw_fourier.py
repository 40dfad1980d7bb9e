"""Hann-windowed (coarse-to-fine) positional encoding, plugin-compatible with reference
`core/nets/human_nerf/embedders/hannw_fourier.py:52-67`:
`get_embedder(multires, iter_val, is_identity)` returns `(embed_fn, out_dim)`; no raw input.
`hann_window_weights` is what `Network.forward` hands to the fused kernels."""
import math

import torch
import torch.nn as nn

from humannerf_b200.config import get_cfg


def hann_window_weights(iter_val, n_freqs: int, kick_in_iter=None, full_band_iter=None) -> torch.Tensor:
    """w_k = (1 - cos(pi * clamp(alpha - k, 0, 1))) / 2 with alpha = m * max(iter - T_s, 0) / (T_e - T_s).
    Evaluated with float32 CPU tensor ops in the reference's order (hannw_fourier.py:24-40)."""
    cfg = get_cfg()
    if kick_in_iter is None:
        kick_in_iter = cfg.non_rigid_motion_mlp.kick_in_iter
    if full_band_iter is None:
        full_band_iter = cfg.non_rigid_motion_mlp.full_band_iter
    it = torch.as_tensor(iter_val).detach().to('cpu', torch.float32).reshape(-1)[:1]
    t_s = torch.tensor(kick_in_iter, dtype=torch.float32)
    t = torch.clamp(it - t_s, min=0.)
    span = full_band_iter - t_s
    alpha = torch.tensor([float(n_freqs)]) if float(span) == 0 else n_freqs * t / span
    k = torch.arange(n_freqs, dtype=torch.float32)
    return ((1. - torch.cos(math.pi * torch.clamp(alpha - k, min=0., max=1.))) / 2.).reshape(n_freqs)


class HannFourierEncoding:
    def __init__(self, n_freqs: int, iter_val):
        self.n_freqs = n_freqs
        self.weights = hann_window_weights(iter_val, n_freqs)
        self.out_dim = 3 * 2 * n_freqs

    def __call__(self, x: torch.Tensor) -> torch.Tensor:
        w = self.weights.to(x.device)
        parts = []
        for k in range(self.n_freqs):
            f = float(2 ** k)
            parts += [w[k] * torch.sin(x * f), w[k] * torch.cos(x * f)]
        return torch.cat(parts, dim=-1)


def get_embedder(multires, iter_val, is_identity=0):
    if is_identity == -1:
        return nn.Identity(), 3
    enc = HannFourierEncoding(multires, iter_val)
    return enc, enc.out_dim
