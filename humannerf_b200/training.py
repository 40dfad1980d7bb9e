"""Loss and train step around the rendering hot path (reference
`core/train/trainers/human_nerf/trainer.py:28-37, 97-175, 194-220`, and the Adam groups of
`core/train/optimizers/human_nerf/optimizer.py:12-43`).

Patch assembly uses the CUDA row scatter (ops.scatter_rows); the perceptual term is the LPIPS(VGG16)
architecture on cuDNN.  The reference's LPIPS needs ImageNet VGG16 weights that cannot be
downloaded offline, so `PerceptualLoss` is seeded-random-initialised:
the LPIPS term is therefore "parity unpinned against real weights"; the MSE term is exact."""
from __future__ import annotations

from typing import Dict, Optional

import torch
import torch.nn as nn

from humannerf_b200 import ops
from humannerf_b200.config import get_cfg


def patch_scatter_index(patch_masks: torch.Tensor) -> torch.Tensor:
    """Flat (N*P*P) destination row of every ray, in ray order: patch i's rays fill its True
    pixels in row-major order (trainer.py:35 `patch_imgs[i, patch_masks[i]] = rgbs[div[i]:div[i+1]]`)."""
    return torch.nonzero(patch_masks.reshape(-1), as_tuple=False)[:, 0]


def unpack_imgs(rgbs: torch.Tensor, patch_masks: torch.Tensor, bgcolor01: torch.Tensor,
                targets: torch.Tensor, div_indices=None, index: Optional[torch.Tensor] = None) -> torch.Tensor:
    """(N_patch, P, P, 3) patches pre-filled with the background colour (trainer.py:28-37)."""
    base = bgcolor01.to(rgbs).expand(targets.shape).contiguous()
    if index is None:
        index = patch_scatter_index(patch_masks)
    return ops.scatter_rows(rgbs, index, base)


class PerceptualLoss(nn.Module):
    """LPIPS v0.1 with a VGG16 trunk (third_parties/lpips/lpips.py:23-129): scale -> five VGG16 stages
    -> unit-normalise channels -> squared difference -> 1x1 non-negative 'lin' -> spatial mean -> sum."""

    def __init__(self, seed: int = 0):
        super().__init__()
        import torchvision
        with torch.random.fork_rng():
            torch.manual_seed(seed)
            feats = torchvision.models.vgg16(weights=None).features
            cuts = (0, 4, 9, 16, 23, 30)
            self.stages = nn.ModuleList([nn.Sequential(*[feats[i] for i in range(a, b)]) for a, b in zip(cuts[:-1], cuts[1:])])
            self.lins = nn.ModuleList([nn.Conv2d(c, 1, 1, bias=False) for c in (64, 128, 256, 512, 512)])
            for l in self.lins:
                l.weight.data.uniform_(0.0, 2.0 / l.weight.shape[1])
        self.register_buffer('shift', torch.tensor([-.030, -.088, -.188])[None, :, None, None])
        self.register_buffer('scale', torch.tensor([.458, .448, .450])[None, :, None, None])
        for p in self.parameters():
            p.requires_grad_(False)
        self.eval()

    def forward(self, x: torch.Tensor, y: torch.Tensor) -> torch.Tensor:
        x, y = (x - self.shift) / self.scale, (y - self.shift) / self.scale
        total = 0.0
        for stage, lin in zip(self.stages, self.lins):
            x, y = stage(x), stage(y)
            nx = x / (torch.sqrt(torch.sum(x ** 2, dim=1, keepdim=True)) + 1e-10)
            ny = y / (torch.sqrt(torch.sum(y ** 2, dim=1, keepdim=True)) + 1e-10)
            total = total + lin((nx - ny) ** 2).mean([2, 3], keepdim=True)
        return total


class RenderLoss(nn.Module):
    """0.2 * MSE + 1.0 * LPIPS over the re-assembled patches (cfg.train.lossweights)."""

    def __init__(self, perceptual: Optional[nn.Module] = None):
        super().__init__()
        self.perceptual = perceptual

    def forward(self, rgb, patch_masks, bgcolor01, targets, index=None) -> Dict[str, torch.Tensor]:
        w = get_cfg().train.lossweights
        imgs = unpack_imgs(rgb, patch_masks, bgcolor01, targets, index=index)
        losses = {}
        if w.get('mse', 0) > 0:
            losses['mse'] = w['mse'] * torch.mean((imgs - targets) ** 2)
        if w.get('l1', 0) > 0:
            losses['l1'] = w['l1'] * torch.mean(torch.abs(imgs - targets))
        if w.get('lpips', 0) > 0 and self.perceptual is not None:
            lp = self.perceptual(imgs.permute(0, 3, 1, 2) * 2. - 1., targets.permute(0, 3, 1, 2) * 2. - 1.)
            losses['lpips'] = w['lpips'] * torch.mean(lp)
        losses['total'] = sum(losses.values())
        return losses


def make_optimizer(network: nn.Module) -> torch.optim.Optimizer:
    """Adam with one group per tensor and per-name learning rates (optimizer.py:12-43): a parameter
    whose name contains `mweight_vol_decoder` / `pose_decoder` / `non_rigid_mlp` uses cfg.train.lr_<name>."""
    cfg = get_cfg()
    custom = [k[3:] for k in cfg.train.keys() if k.startswith('lr_')]
    groups = []
    for name, p in network.named_parameters():
        if not p.requires_grad:
            continue
        lr_name = next((c for c in custom if c in name), None)
        if lr_name is not None:
            groups.append({'params': [p], 'lr': cfg.train[f'lr_{lr_name}'], 'name': lr_name})
        else:
            groups.append({'params': [p], 'name': name})
    return torch.optim.Adam(groups, lr=cfg.train.lr, betas=(0.9, 0.999))


def update_lr(optimizer: torch.optim.Optimizer, iter_step: int) -> None:
    """lr = base * 0.1 ** (iter / (lrate_decay * 1000))  (lr_updaters/exp_decay.py:7-16)."""
    cfg = get_cfg()
    decay = 0.1 ** (iter_step / (cfg.train.lrate_decay * 1000))
    for g in optimizer.param_groups:
        base = cfg.train[f"lr_{g['name']}"] if f"lr_{g['name']}" in cfg.train else cfg.train.lr
        g['lr'] = base * decay
