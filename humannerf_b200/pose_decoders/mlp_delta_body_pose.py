"""Body-pose refiner, plugin-compatible with reference
`core/nets/human_nerf/pose_decoders/mlp_delta_body_pose.py:7-41`: 69-d pose -> MLP -> 23
axis-angle corrections -> rotation matrices.  Batch 1, stays in PyTorch."""
import torch.nn as nn

from humannerf_b200.config import get_cfg
from humannerf_b200.nn_util import RodriguesModule, init_sequence


class BodyPoseRefiner(nn.Module):
    def __init__(self, embedding_size=69, mlp_width=256, mlp_depth=4, **_):
        super().__init__()
        layers = [nn.Linear(embedding_size, mlp_width), nn.ReLU()]
        for _i in range(mlp_depth - 1):
            layers += [nn.Linear(mlp_width, mlp_width), nn.ReLU()]
        self.total_bones = get_cfg().total_bones - 1
        layers += [nn.Linear(mlp_width, 3 * self.total_bones)]
        self.block_mlps = nn.Sequential(*layers)
        init_sequence(self.block_mlps)
        last = self.block_mlps[-1]
        last.weight.data.uniform_(-1e-5, 1e-5)
        last.bias.data.zero_()
        self.rodriguez = RodriguesModule()

    def forward(self, pose_input):
        rvec = self.block_mlps(pose_input).view(-1, 3)
        Rs = self.rodriguez(rvec).view(-1, self.total_bones, 3, 3)
        return {'Rs': Rs, 'rvec': rvec.view(-1, self.total_bones, 3)}
