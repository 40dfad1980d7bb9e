"""Motion-weight volume decoder, plugin-compatible with reference
`core/nets/human_nerf/mweight_vol_decoders/deconv_vol_decoder.py:8-33`:
softmax(decoder(const_embedding) + log(prior)) over the 25 channels.  Runs once per forward on
cuDNN (dense transposed convolutions); SURVEY.md section 8 keeps it in PyTorch."""
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
        # set by humannerf_b200.Network: True on the bf16 tensor-core path (2e-2 parity bound), where the
        # TF32 transposed convolutions are ~0.4 ms per training step faster than cuDNN's fp32 engine
        self.allow_tf32 = False

    def forward(self, motion_weights_priors, **_):
        # on the fp32 parity path TF32 is off: its 10-bit mantissa moves the sampled blend weights by 1e-4
        # and breaks the 1e-4 output parity with the reference
        with torch.backends.cudnn.flags(enabled=True, allow_tf32=bool(self.allow_tf32)):
            logits = self.decoder(self.const_embedding[None, ...])
        return F.softmax(logits + torch.log(motion_weights_priors), dim=1)
