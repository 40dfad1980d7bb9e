"""Motion-weight volume decoder, plugin-compatible with reference
`core/nets/human_nerf/mweight_vol_decoders/deconv_vol_decoder.py:8-33`:
softmax(decoder(const_embedding) + log(prior)) over the 25 channels.  On a CUDA device the whole module - linear
layer, the five transposed convolutions, LeakyReLUs, softmax with the prior, and their backward - runs on the
hand-written kernels of csrc/vol_decoder.cu (`ops.vol_decoder`: fp32 storage, tensor-core products, deterministic).
`cfg.b200.decoder`: 'hn' / 'auto' (default) = those kernels (fp32 parity path: 3xTF32 products; bf16 path: TF32 products
like cuDNN's TF32 engines), 'cudnn' = PyTorch / cuDNN (kept for A/B: 6.67 against 6.71 ms per training step).
CPU tensors (construction-time tests, the oracle arm) always take the PyTorch path."""
import torch
import torch.nn as nn
import torch.nn.functional as F

from humannerf_b200.nn_util import ConvDecoder3D


class MotionWeightVolumeDecoder(nn.Module):
    def __init__(self, embedding_size=256, volume_size=32, total_bones=24):
        super().__init__()
        self.total_bones = total_bones
        self.volume_size = volume_size
        self.const_embedding = nn.Parameter(torch.randn(embedding_size), requires_grad=True)
        self.decoder = ConvDecoder3D(embedding_size=embedding_size, volume_size=volume_size,
                                     voxel_channels=total_bones + 1)
        # set by humannerf_b200.Network: True on the bf16 tensor-core path (2e-2 parity bound): plain TF32 products
        self.allow_tf32 = False

    def _use_kernels(self, prior) -> bool:
        from humannerf_b200.config import get_cfg
        mode = str(get_cfg().get('b200', {}).get('decoder', 'auto')).lower()
        return prior.is_cuda and mode != 'cudnn'

    def forward(self, motion_weights_priors, **_):
        if self._use_kernels(motion_weights_priors):
            from humannerf_b200 import ops
            convs = [m for m in self.decoder.block_conv if isinstance(m, nn.ConvTranspose3d)]
            lin = self.decoder.block_mlp[0]
            prior = motion_weights_priors[0] if motion_weights_priors.dim() == 5 else motion_weights_priors
            from humannerf_b200 import _lib
            # allow_tf32 (set by Network: True on the bf16 tensor-core path) selects plain TF32 products; the fp32
            # parity path multiplies as 3xTF32 (fp32-level accuracy)
            vol = ops.vol_decoder(self.const_embedding, lin.weight, lin.bias, prior, [c.weight for c in convs],
                                  [c.bias for c in convs],
                                  precision=_lib.PRECISION_BF16 if self.allow_tf32 else _lib.PRECISION_FP32)
            return vol[None, ...]
        # on the fp32 parity path TF32 is off: its 10-bit mantissa moves the sampled blend weights by 1e-4
        # and breaks the 1e-4 output parity with the reference
        with torch.backends.cudnn.flags(enabled=True, allow_tf32=bool(self.allow_tf32)):
            logits = self.decoder(self.const_embedding[None, ...])
        return F.softmax(logits + torch.log(motion_weights_priors), dim=1)
