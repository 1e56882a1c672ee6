"""Random-init stand-in for the DINOv2 ViT-L/14 backbone of the DINO branch (reference dataset.py:61-80, `DINOV2`).

The backbone is NOT part of the hot path this repository rebuilds: the reference downloads its weights with
`torch.hub.load('facebookresearch/dinov2', 'dinov2_vitl14')`, which no offline box can do, and the path takes the [N,1024]
key-point descriptors as an input.  This module exists so that the frame loop can be driven from an RGB image end to end
(SURVEY.md section 8f, rank 3): a plain-PyTorch encoder with the ViT-L/14 shape (patch 14, width 1024, 24 blocks, 16 heads,
MLP 4096, LayerScale, class token, learned position table interpolated to the patch grid) and seeded random weights,
followed by the reference's own post-processing -- `x_norm_patchtokens` viewed as [1,C,h,w] and sampled at the key-points by
`cppf2_b200.cloud.interpolate_features` (the CUDA kernel pinned on tests/golden/interp_features.npz).  PyTorch modules are
library code here, like the real backbone is for the reference; nothing in it is tuned.

    net = DINOV2StandIn(stride=4).cuda().eval()          # dataset.py:61-66
    feats = net(rgb, pts)                                 # rgb [3,H,W] in [0,1], pts [n,2] (x, y) pixels -> [n,1024], unit rows
    poses = estimator.estimate_frame(depth, masks, cats, K, desc_fn=net.desc_fn(rgb))
"""
from __future__ import annotations

import math
from typing import Callable

import torch
import torch.nn as nn
import torch.nn.functional as F

IMAGENET_MEAN = (0.485, 0.456, 0.406)      # dataset.py:73
IMAGENET_STD = (0.229, 0.224, 0.225)


class _Block(nn.Module):
    """Pre-norm transformer block with LayerScale (the DINOv2 block: norm -> attention -> scale, norm -> MLP -> scale)."""

    def __init__(self, dim: int, heads: int, mlp_ratio: float):
        super().__init__()
        self.norm1 = nn.LayerNorm(dim, eps=1e-6)
        self.qkv = nn.Linear(dim, 3 * dim)
        self.proj = nn.Linear(dim, dim)
        self.ls1 = nn.Parameter(torch.ones(dim))
        self.norm2 = nn.LayerNorm(dim, eps=1e-6)
        self.fc1 = nn.Linear(dim, int(dim * mlp_ratio))
        self.fc2 = nn.Linear(int(dim * mlp_ratio), dim)
        self.ls2 = nn.Parameter(torch.ones(dim))
        self.heads = heads

    def forward(self, x: torch.Tensor) -> torch.Tensor:
        b, n, c = x.shape
        qkv = self.qkv(self.norm1(x)).reshape(b, n, 3, self.heads, c // self.heads).permute(2, 0, 3, 1, 4)
        a = F.scaled_dot_product_attention(qkv[0], qkv[1], qkv[2])
        x = x + self.ls1 * self.proj(a.transpose(1, 2).reshape(b, n, c))
        return x + self.ls2 * self.fc2(F.gelu(self.fc1(self.norm2(x))))


class DINOV2StandIn(nn.Module):
    """Same call as the reference's `DINOV2` (dataset.py:61-80): forward(rgb [3,H,W], pts [n,2]) -> [n, width] unit rows."""

    def __init__(self, stride: int = 4, width: int = 1024, depth: int = 24, heads: int = 16, mlp_ratio: float = 4.0,
                 patch: int = 14, table: int = 37, seed: int = 0):
        super().__init__()
        self.stride, self.patch, self.width, self.table = stride, patch, width, table
        gen = torch.Generator().manual_seed(seed)
        self.patch_embed = nn.Conv2d(3, width, patch, patch)
        self.cls_token = nn.Parameter(torch.zeros(1, 1, width))
        self.pos_embed = nn.Parameter(torch.zeros(1, 1 + table * table, width))      # 518 / 14 = 37 tokens a side
        self.blocks = nn.ModuleList([_Block(width, heads, mlp_ratio) for _ in range(depth)])
        self.norm = nn.LayerNorm(width, eps=1e-6)
        with torch.no_grad():                            # seeded: the same stand-in on every rank and every run
            for name, p in self.named_parameters():     # matrices and tables drawn from `gen`, biases zero, norms / LayerScale one
                if p.dim() > 1:
                    p.copy_(torch.randn(p.shape, generator=gen) * (0.02 if p is not self.patch_embed.weight else 1.0 / math.sqrt(3 * patch * patch)))
                elif name.endswith("bias"):
                    p.zero_()
            self.cls_token.copy_(torch.randn(self.cls_token.shape, generator=gen) * 1e-6)

    def _positions(self, h: int, w: int) -> torch.Tensor:
        """The learned table resampled to the h x w patch grid (bicubic, like the hub model does for other input sizes)."""
        cls, grid = self.pos_embed[:, :1], self.pos_embed[:, 1:]
        if (h, w) != (self.table, self.table):
            grid = grid.reshape(1, self.table, self.table, self.width).permute(0, 3, 1, 2)
            grid = F.interpolate(grid, size=(h, w), mode="bicubic", align_corners=False).permute(0, 2, 3, 1).reshape(1, h * w, self.width)
        return torch.cat([cls, grid], 1)

    def patch_tokens(self, rgb: torch.Tensor) -> torch.Tensor:
        """rgb [3,H,W] in [0,1] -> normalised patch tokens as the reference views them, [1, width, H // stride, W // stride]
        (dataset.py:69-78: resize so that one 14-pixel patch covers `stride` input pixels, ImageNet normalisation,
        `forward_features(...)['x_norm_patchtokens']`)."""
        ph, pw = rgb.shape[-2] // self.stride, rgb.shape[-1] // self.stride
        x = F.interpolate(rgb[None].float(), size=(ph * self.patch, pw * self.patch), mode="bilinear", antialias=True, align_corners=False)
        mean = torch.tensor(IMAGENET_MEAN, device=x.device).view(1, 3, 1, 1)
        std = torch.tensor(IMAGENET_STD, device=x.device).view(1, 3, 1, 1)
        x = self.patch_embed((x - mean) / std).flatten(2).transpose(1, 2)                # [1, ph*pw, width]
        x = torch.cat([self.cls_token.expand(1, -1, -1), x], 1) + self._positions(ph, pw)
        for blk in self.blocks:
            x = blk(x)
        tokens = self.norm(x)[:, 1:]                                                      # x_norm_patchtokens
        return tokens.reshape(1, ph, pw, self.width).permute(0, 3, 1, 2)                 # a permuted view, like dataset.py:77

    @torch.no_grad()
    def forward(self, rgb: torch.Tensor, pts: torch.Tensor) -> torch.Tensor:
        from .cloud import interpolate_features
        raw = self.patch_tokens(rgb)
        return interpolate_features(raw, pts[None], strides=self.stride, normalize=True)[0].T

    def desc_fn(self, rgb: torch.Tensor) -> Callable:
        """The `desc_fn(i, pix)` of `PoseEstimator.estimate_frame`: the tokens of the whole frame are computed once and every
        detection samples them at its kept pixels (pix = row * W + col).  The reference runs the backbone once per detection
        on a crop around it (eval.py:168-183, 203-205); that call is `forward(rgb_local, kp_local)` above -- one pass per
        frame is the cheaper schedule a stand-in can afford to offer next to it."""
        from .cloud import interpolate_features
        with torch.no_grad():
            raw = self.patch_tokens(rgb)
        w = rgb.shape[-1]

        def fn(_i: int, pix: torch.Tensor) -> torch.Tensor:
            p = pix.to(torch.int64)
            pts = torch.stack([(p % w).float(), torch.div(p, w, rounding_mode="floor").float()], -1)      # (x, y)
            return interpolate_features(raw, pts[None], strides=self.stride, normalize=True)[0].T.contiguous()
        return fn
