"""Non-rigid motion MLP container, plugin-compatible with reference
`core/nets/human_nerf/non_rigid_motion_mlps/mlp_offset.py:8-114` (single head).

Parameters live under the reference's names (`block_mlps.{0,2,..,12}`); `Network.forward`
feeds them to the fused CUDA field kernels.  Direct calls use torch ops (compatibility only)."""
import torch
import torch.nn as nn

from humannerf_b200.nn_util import init_sequence


class NonRigidMotionMLP(nn.Module):
    def __init__(self, pos_embed_size=3, condition_code_size=69, mlp_width=128, mlp_depth=6, mlp_depth_plus=0,
                 skips=None, multihead_enable=False, multihead_depth=1, multihead_num=4, last_linear_scale=1,
                 **kwargs):
        super().__init__()
        if multihead_enable or last_linear_scale != 1:
            raise NotImplementedError('humannerf_b200.NonRigidMotionMLP implements the single-head MLP only')
        self.skips = [4] if skips is None else list(skips)
        self.pos_embed_size, self.condition_code_size = pos_embed_size, condition_code_size
        layers = [nn.Linear(pos_embed_size + condition_code_size, mlp_width), nn.ReLU()]
        self.layers_to_cat_inputs = []
        for i in range(1, mlp_depth + mlp_depth_plus):
            if i in self.skips:
                self.layers_to_cat_inputs.append(len(layers))
                layers += [nn.Linear(mlp_width + pos_embed_size, mlp_width), nn.ReLU()]
            else:
                layers += [nn.Linear(mlp_width, mlp_width), nn.ReLU()]
        layers += [nn.Linear(mlp_width, 3)]
        self.block_mlps = nn.ModuleList(layers)
        init_sequence(self.block_mlps)
        # offsets start at (almost) zero
        last = self.block_mlps[-1]
        last.weight.data.uniform_(-1e-5, 1e-5)
        last.bias.data.zero_()

    def linear_layers(self):
        return [m for m in self.block_mlps if isinstance(m, nn.Linear)]

    def skip_layer_index(self) -> int:
        return self.layers_to_cat_inputs[0] // 2 if self.layers_to_cat_inputs else -1

    def forward(self, pos_embed, pos_xyz, condition_code, **_):
        h = torch.cat([condition_code.expand(pos_embed.shape[0], -1), pos_embed], dim=-1)
        for i, layer in enumerate(self.block_mlps):
            if i in self.layers_to_cat_inputs:
                h = torch.cat([h, pos_embed], dim=-1)
            h = layer(h)
        return {'xyz': pos_xyz + h, 'offsets': h}
