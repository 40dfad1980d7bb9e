"""Torch-CPU emulation of the numerics of the bf16 tensor-core field path (test infrastructure).

Same graph as `oracle.humannerf_oracle.non_rigid_mlp` / `canonical_mlp` (reference
`mlp_offset.py:74-114`, `mlp_rgb_sigma.py:132-206`) with the rounding points of
`humannerf_b200/csrc/mlp_tc_2sm.cu` / `mlp_tc_bwd_2sm.cu`:

* weights, biases (the first non-rigid bias with the condition code folded in) and encodings are
  rounded to bf16, products accumulate in fp32;
* hidden activations are stored as bf16 after the ReLU;
* in the backward pass every pre-activation gradient dZ is rounded to bf16 before it is used as a
  GEMM operand (data gradient and weight gradient); encoding gradients stay fp32.

Rounding is straight-through for autograd, so `.backward()` on the emulated outputs yields what the
CUDA kernels are supposed to compute up to fp32 summation order.
"""
import torch
import torch.nn.functional as F

from oracle import humannerf_oracle as orc


def _q(x):
    return x.to(torch.bfloat16).to(torch.float32)


class _RoundSTE(torch.autograd.Function):
    """bf16 rounding forward, identity backward."""
    @staticmethod
    def forward(ctx, x):
        return _q(x)

    @staticmethod
    def backward(ctx, g):
        return g


class _RoundGrad(torch.autograd.Function):
    """identity forward, bf16 rounding of the gradient."""
    @staticmethod
    def forward(ctx, x):
        return x.view_as(x)

    @staticmethod
    def backward(ctx, g):
        return _q(g)


def rq(x):
    return _RoundSTE.apply(x)


def gq(x):
    return _RoundGrad.apply(x)


def non_rigid_mlp(sd, pe, pos, cond, skip_at=8, prefix='non_rigid_mlp.module.block_mlps'):
    ids = sorted({int(k[len(prefix) + 1:].split('.')[0]) for k in sd if k.startswith(prefix + '.')})
    peq = rq(pe)
    n_cond = cond.shape[-1]
    h = None
    for i in ids:
        W, b = sd[f'{prefix}.{i}.weight'], sd[f'{prefix}.{i}.bias']
        if i == ids[0]:
            bias = rq(b + W[:, :n_cond] @ cond.reshape(-1))
            z = F.linear(peq, rq(W[:, n_cond:])) + bias
        elif i == skip_at:
            z = F.linear(torch.cat([h, peq], dim=-1), rq(W)) + rq(b)
        else:
            z = F.linear(h, rq(W)) + rq(b)
        z = gq(z)
        if i != ids[-1]:
            h = rq(torch.relu(z))
    return pos + z, z


def canonical_mlp(sd, pe, skip_at=10, prefix='cnl_mlp.module'):
    trunk = f'{prefix}.pts_linears'
    ids = sorted({int(k[len(trunk) + 1:].split('.')[0]) for k in sd if k.startswith(trunk + '.')})
    peq = rq(pe)
    h = peq
    for i in ids:
        if i == skip_at:
            h = torch.cat([peq, h], dim=-1)
        z = gq(F.linear(h, rq(sd[f'{trunk}.{i}.weight'])) + rq(sd[f'{trunk}.{i}.bias']))
        h = rq(torch.relu(z))
    W, b = sd[f'{prefix}.output_linear.0.weight'], sd[f'{prefix}.output_linear.0.bias']
    return gq(F.linear(h, rq(W)) + rq(b))


def field(sd, x, cond, hann, ignore_non_rigid=False, nr_freqs=6, cn_freqs=10):
    """(raw, xyz, offsets) of the bf16 path for canonical positions `x` (P,3)."""
    if ignore_non_rigid:
        xyz, off = x, torch.zeros_like(x)
    else:
        xyz, off = non_rigid_mlp(sd, orc.fourier_pe(x, nr_freqs, hann, False), x, cond)
    raw = canonical_mlp(sd, orc.fourier_pe(xyz, cn_freqs))
    return raw, xyz, off
