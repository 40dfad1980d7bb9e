"""Per-frame building blocks kept in PyTorch (tiny, batch 1): weight initialisation,
the volume decoder trunk, Rodrigues and the motion-basis computer.

They mirror the behaviour of reference `core/utils/network_util.py` (initseq :274-291,
ConvDecoder3D :12-50, RodriguesModule :57-83, MotionBasisComputer :97-156) so that a
network built under the same `torch.manual_seed` gets bit-identical initial weights
(checked by tests/test_host.py against golden statistics of the reference)."""
from __future__ import annotations

import math

import torch
import torch.nn as nn

SMPL_PARENT = (-1, 0, 0, 0, 1, 2, 3, 4, 5, 6, 7, 8, 9, 9, 9, 12, 13, 14, 16, 17, 18, 19, 20, 21)


def _xavier_std(m: nn.Module, gain: float):
    if isinstance(m, nn.Linear):
        return gain * math.sqrt(2.0 / (m.in_features + m.out_features))
    if isinstance(m, nn.ConvTranspose3d):
        k = (m.kernel_size[0] * m.kernel_size[1] * m.kernel_size[2]
             // m.stride[0] // m.stride[1] // m.stride[2])
        return gain * math.sqrt(2.0 / ((m.in_channels + m.out_channels) * k))
    return None


def init_module(m: nn.Module, gain: float = 1.0) -> None:
    """Xavier-uniform weights, zero bias; stride-2 transposed convolutions get the same
    value in all 8 sub-positions of each 2x2x2 output block."""
    std = _xavier_std(m, gain)
    if std is None:
        return
    bound = std * math.sqrt(3.0)
    m.weight.data.uniform_(-bound, bound)
    if getattr(m, 'bias', None) is not None:
        m.bias.data.zero_()
    if isinstance(m, nn.ConvTranspose3d):
        w = m.weight.data
        base = w[:, :, 0::2, 0::2, 0::2]
        for a in (0, 1):
            for b in (0, 1):
                for c in (0, 1):
                    if a or b or c:
                        w[:, :, a::2, b::2, c::2] = base


def init_sequence(seq) -> None:
    """Gain follows the activation that comes right after each layer; the last module uses gain 1."""
    mods = list(seq)
    for cur, nxt in zip(mods[:-1], mods[1:]):
        if isinstance(nxt, nn.ReLU):
            init_module(cur, nn.init.calculate_gain('relu'))
        elif isinstance(nxt, nn.LeakyReLU):
            init_module(cur, nn.init.calculate_gain('leaky_relu', nxt.negative_slope))
        else:
            init_module(cur)
    init_module(mods[-1])


class _FirstDeconv(torch.autograd.Function):
    """ConvTranspose3d(k4, s2, p1) on a 1^3 input as a product with the central 2^3 taps (see ConvDecoder3D.forward).
    Written as a Function because autograd's own backward of `weight[:, :, 1:3, 1:3, 1:3]` builds the full-size
    gradient in contiguous NCDHW strides and AccumulateGrad then re-strides all 33.5 M elements to the parameter's
    NDHWC layout (a 268 MB copy, 0.14 ms per training step); here the gradient is allocated in the parameter's own
    strides (the untouched taps are exactly zero, as they are in the reference's full transposed convolution)."""

    @staticmethod
    def forward(ctx, h, weight, bias):
        taps = weight[:, :, 1:3, 1:3, 1:3]
        ctx.save_for_backward(h, weight)
        return torch.einsum('bc,cozyx->bozyx', h, taps) + bias.view(1, -1, 1, 1, 1)

    @staticmethod
    def backward(ctx, g):
        h, weight = ctx.saved_tensors
        d_h = d_w = d_b = None
        if ctx.needs_input_grad[0]:
            d_h = torch.einsum('bozyx,cozyx->bc', g, weight[:, :, 1:3, 1:3, 1:3])
        if ctx.needs_input_grad[1]:
            d_w = torch.zeros_like(weight)                    # preserve_format: same (NDHWC) strides as the parameter
            d_w[:, :, 1:3, 1:3, 1:3] = torch.einsum('bc,bozyx->cozyx', h, g)
        if ctx.needs_input_grad[2]:
            d_b = g.sum(dim=(0, 2, 3, 4))
        return d_h, d_w, d_b


class ConvDecoder3D(nn.Module):
    """embedding -> Linear(., 1024) + LeakyReLU(0.2) -> (1024,1,1,1) -> log2(V) transposed
    convolutions (k4, s2, p1) halving the channels every other layer."""

    def __init__(self, embedding_size=256, volume_size=128, voxel_channels=4):
        super().__init__()
        self.block_mlp = nn.Sequential(nn.Linear(embedding_size, 1024), nn.LeakyReLU(0.2))
        layers = []
        cin, cout = 1024, 512
        for _ in range(int(math.log2(volume_size)) - 1):
            layers += [nn.ConvTranspose3d(cin, cout, 4, 2, 1), nn.LeakyReLU(0.2)]
            if cin == cout:
                cout = cin // 2
            else:
                cin = cout
        layers.append(nn.ConvTranspose3d(cin, voxel_channels, 4, 2, 1))
        self.block_conv = nn.Sequential(*layers)
        init_sequence(self.block_mlp)
        init_sequence(self.block_conv)
        # same values and shapes as the reference, NDHWC strides: cuDNN's transposed-convolution engines work in
        # NDHWC and would otherwise transpose the 63.6 M weights (and their gradients) on every call, ~0.5 ms per step
        self.block_conv.to(memory_format=torch.channels_last_3d)

    def forward(self, embedding):
        h = self.block_mlp(embedding)                         # (B, 1024): a 1 x 1 x 1 volume
        # First transposed convolution (k4, s2, p1) on a 1^3 input: output voxel o takes tap k = o + 1, so only the
        # central 2^3 of the 4^3 taps are ever used - a (B, 1024) x (1024, 512 * 8) product.  cuDNN would transpose
        # and stream all 33.5 M weights three times per training step; the other taps get exactly zero gradient.
        conv0 = self.block_conv[0]
        assert conv0.kernel_size == (4, 4, 4) and conv0.stride == (2, 2, 2) and conv0.padding == (1, 1, 1)
        x = _FirstDeconv.apply(h, conv0.weight, conv0.bias)
        return self.block_conv[1:](x)


class RodriguesModule(nn.Module):
    def forward(self, rvec):
        theta = torch.sqrt(1e-5 + torch.sum(rvec ** 2, dim=1))
        u = rvec / theta[:, None]
        c, s = torch.cos(theta), torch.sin(theta)
        x, y, z = u[:, 0], u[:, 1], u[:, 2]
        return torch.stack((
            x * x + (1. - x * x) * c, x * y * (1. - c) - z * s, x * z * (1. - c) + y * s,
            x * y * (1. - c) + z * s, y * y + (1. - y * y) * c, y * z * (1. - c) - x * s,
            x * z * (1. - c) - y * s, y * z * (1. - c) + x * s, z * z + (1. - z * z) * c), dim=1).view(-1, 3, 3)


class MotionBasisComputer(nn.Module):
    """Forward kinematics over the SMPL tree and canonical @ inverse(posed) per joint."""

    def __init__(self, total_bones=24):
        super().__init__()
        self.total_bones = total_bones

    def forward(self, dst_Rs, dst_Ts, cnl_gtfms):
        B, J = dst_Rs.shape[:2]
        assert J == self.total_bones
        local = torch.zeros(B, J, 4, 4, dtype=dst_Rs.dtype, device=dst_Rs.device)
        local[:, :, :3, :3] = dst_Rs
        local[:, :, :3, 3] = dst_Ts
        local[:, :, 3, 3] = 1.0
        glob = [local[:, 0]]
        for i in range(1, J):
            glob.append(torch.matmul(glob[SMPL_PARENT[i]], local[:, i]))
        glob = torch.stack(glob, dim=1).view(-1, 4, 4)
        # The reference calls torch.inverse (network_util.py:148).  linalg.inv_ex is the same LU
        # factorisation without the host-side check of the `info` flag, i.e. without the device->host
        # synchronisation that would break CUDA-graph capture of the train step.  (An analytic rigid
        # inverse [R^T, -R^T t] was tried: it differs from the LU result by ~1e-6, which the 2^9 encoding
        # frequency amplifies beyond the 1e-4 parity bound of the fp32 path.)
        inv = torch.linalg.inv_ex(glob, check_errors=False).inverse
        f = torch.matmul(cnl_gtfms.reshape(-1, 4, 4), inv).view(B, J, 4, 4)
        return f[:, :, :3, :3], f[:, :, :3, 3]
