"""Canonical NeRF MLP container, plugin-compatible with reference
`core/nets/human_nerf/canonical_mlps/mlp_rgb_sigma.py:13-206` (default branch: no view
directions, pose_color 'wo', single head, no input encoder, no condition code).

The module owns the parameters under the reference's names (`pts_linears.{0,2,..}`,
`output_linear.0`); `Network.forward` feeds them to the fused CUDA field kernels.  Calling the
module directly evaluates the same layers with torch ops (compatibility only)."""
import torch
import torch.nn as nn

from humannerf_b200.nn_util import init_sequence


class CanonicalMLP(nn.Module):
    def __init__(self, mlp_depth=8, mlp_width=256, input_ch=3, skips=None, view_dir=False, input_ch_dir=3,
                 pose_ch=69, pose_color='wo', multihead_enable=False, multihead_depth=1, multihead_num=4,
                 mlp_depth_plus=0, last_linear_scale=1, rgb_dynamic_features_ch=0, **kwargs):
        super().__init__()
        if view_dir or pose_color != 'wo' or multihead_enable or rgb_dynamic_features_ch or last_linear_scale != 1:
            raise NotImplementedError(
                'humannerf_b200.CanonicalMLP implements the default branch only (view_dir=False, '
                "pose_color='wo', single head, no rgb history, last_linear_scale=1)")
        skips = [4] if skips is None else list(skips)
        self.mlp_width, self.input_ch = mlp_width, input_ch
        layers = [nn.Linear(input_ch, mlp_width), nn.ReLU()]
        self.layers_to_cat_input = []
        for i in range(mlp_depth + mlp_depth_plus - 1):
            if i in skips:
                self.layers_to_cat_input.append(len(layers))
                layers += [nn.Linear(mlp_width + input_ch, mlp_width), nn.ReLU()]
            else:
                layers += [nn.Linear(mlp_width, mlp_width), nn.ReLU()]
        self.pts_linears = nn.ModuleList(layers)
        init_sequence(self.pts_linears)
        self.output_linear = nn.Sequential(nn.Linear(mlp_width, 4))
        init_sequence(self.output_linear)

    # -- description consumed by Network ------------------------------------------------
    def linear_layers(self):
        return [m for m in self.pts_linears if isinstance(m, nn.Linear)] + [self.output_linear[0]]

    def skip_layer_index(self) -> int:
        """Index (among the Linear layers) of the layer that re-reads the encoding; -1 if none."""
        return self.layers_to_cat_input[0] // 2 if self.layers_to_cat_input else -1

    def forward(self, pos_embed, **_):
        h = pos_embed
        for i, layer in enumerate(self.pts_linears):
            if i in self.layers_to_cat_input:
                h = torch.cat([pos_embed, h], dim=-1)
            h = layer(h)
        return self.output_linear(h)
