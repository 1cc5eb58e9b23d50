"""Construction-time ViT surgery of the reference (``src/depth_pro/network/vit.py:51-123``, driven by
``vit_factory.py:104-110``), as pure functions over state-dict tensors.

The reference builds every encoder as timm's ``vit_large_patch14_dinov2`` (patch 14, image 518: a 16x16 patch conv it is
not, and a 37x37 + cls position table) and rewrites it in place to the ``dinov2l16_384`` preset: ``resize_patch_embed``
resamples the patch-embedding conv 14x14 -> 16x16 (bicubic, rescaled by (14/16)^2 so a patch's response keeps its
magnitude) and ``resize_vit`` resamples the position table 37x37 -> 24x24 (timm ``resample_abs_pos_embed``: bicubic,
antialiased, prefix token kept).  A published ``depth_pro.pt`` already holds the post-surgery tensors, so the engine never
needs this at inference time; it exists so that a RAW timm / DINOv2-L state dict (the only other place such weights come
from) can be brought to the shapes the engine ingests.  One-time host work on the CPU with torch ops -- not part of the
per-frame path.
"""

from __future__ import annotations

import math
from typing import Dict, Mapping, Tuple

import torch
import torch.nn.functional as F

__all__ = ["resize_patch_embed_weight", "resample_pos_embed", "convert_timm_vit_state_dict"]


def resize_patch_embed_weight(weight: torch.Tensor, new_patch_size: Tuple[int, int] = (16, 16)) -> torch.Tensor:
    """``resize_patch_embed`` (vit.py:70-123) on the conv weight (O, 3, h, w); the bias is unchanged."""
    _, _, h, w = weight.shape
    if (h, w) == tuple(new_patch_size):
        return weight
    out = F.interpolate(weight, size=[new_patch_size[0], new_patch_size[1]], mode="bicubic", align_corners=False)
    return out * (h / new_patch_size[0]) * (w / new_patch_size[1])


def resample_pos_embed(pos_embed: torch.Tensor, grid_size: Tuple[int, int] = (24, 24), num_prefix_tokens: int = 1
                       ) -> torch.Tensor:
    """``resize_vit`` (vit.py:51-67) = timm ``resample_abs_pos_embed(pos_embed, grid_size, num_prefix_tokens)`` with its
    defaults (bicubic, antialias=True): (1, P + h*w, C) -> (1, P + H*W, C)."""
    n_old = pos_embed.shape[1]
    if grid_size[0] * grid_size[1] + num_prefix_tokens == n_old and grid_size[0] == grid_size[1]:
        return pos_embed
    hw = int(math.sqrt(n_old - num_prefix_tokens))
    prefix, grid = pos_embed[:, :num_prefix_tokens], pos_embed[:, num_prefix_tokens:]
    c = grid.shape[-1]
    g = grid.float().reshape(1, hw, hw, c).permute(0, 3, 1, 2)
    g = F.interpolate(g, size=grid_size, mode="bicubic", antialias=True)
    g = g.permute(0, 2, 3, 1).reshape(1, -1, c).to(pos_embed.dtype)
    return torch.cat([prefix, g], dim=1) if num_prefix_tokens else g


def convert_timm_vit_state_dict(sd: Mapping[str, torch.Tensor], prefix: str = "", img_size: int = 384,
                                patch_size: int = 16) -> Dict[str, torch.Tensor]:
    """A raw timm ``vit_large_patch14_dinov2`` state dict -> the ``dinov2l16_384`` tensors the reference's
    ``create_vit`` leaves behind (vit_factory.py:97-110), under ``prefix`` (e.g. ``"encoder.patch_encoder."``)."""
    out = {}
    for k, v in sd.items():
        if k == "patch_embed.proj.weight":
            v = resize_patch_embed_weight(v, (patch_size, patch_size))
        elif k == "pos_embed":
            v = resample_pos_embed(v, (img_size // patch_size, img_size // patch_size), 1)
        out[prefix + k] = v
    return out
