"""NeRF positional encoding, plugin-compatible with reference
`core/nets/human_nerf/embedders/fourier.py:41-55`: `get_embedder(multires, i)` returns
`(embed_fn, out_dim)`.  `Network.forward` never calls `embed_fn` - the encoding is fused into
the canonical-field kernels (ops.canonical_field); the callable is kept for callers that use
the embedder on its own and is written with plain torch ops."""
import torch
import torch.nn as nn


class FourierEncoding:
    def __init__(self, n_freqs: int, include_input: bool = True, input_dims: int = 3):
        self.n_freqs = n_freqs
        self.include_input = include_input
        self.out_dim = input_dims * (2 * n_freqs + (1 if include_input else 0))

    def __call__(self, x: torch.Tensor) -> torch.Tensor:
        parts = [x] if self.include_input else []
        for k in range(self.n_freqs):
            f = float(2 ** k)
            parts += [torch.sin(x * f), torch.cos(x * f)]
        return torch.cat(parts, dim=-1)


def get_embedder(multires, i=0, include_input=True):
    if i == -1:
        return nn.Identity(), 3
    enc = FourierEncoding(multires, include_input)
    return enc, enc.out_dim
