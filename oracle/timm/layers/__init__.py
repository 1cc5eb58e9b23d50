"""TEST INFRASTRUCTURE ONLY — `timm.layers.resample_abs_pos_embed` restated (see ../__init__.py).

Used by the reference at `src/depth_pro/network/vit.py:5,58-65` (construction-time
resample 37x37 -> 24x24) and by the shim's `_pos_embed` (early-return no-op at 384x384).
"""

import math

import torch
import torch.nn.functional as F


def resample_abs_pos_embed(posemb, new_size, old_size=None, num_prefix_tokens=1,
                           interpolation="bicubic", antialias=True, verbose=False):
    num_pos_tokens = posemb.shape[1]
    num_new_tokens = new_size[0] * new_size[1] + num_prefix_tokens
    if num_new_tokens == num_pos_tokens and new_size[0] == new_size[1]:
        return posemb

    if old_size is None:
        hw = int(math.sqrt(num_pos_tokens - num_prefix_tokens))
        old_size = hw, hw

    if num_prefix_tokens:
        posemb_prefix, posemb = posemb[:, :num_prefix_tokens], posemb[:, num_prefix_tokens:]
    else:
        posemb_prefix = None

    embed_dim = posemb.shape[-1]
    orig_dtype = posemb.dtype
    posemb = posemb.float()
    posemb = posemb.reshape(1, old_size[0], old_size[1], -1).permute(0, 3, 1, 2)
    posemb = F.interpolate(posemb, size=new_size, mode=interpolation, antialias=antialias)
    posemb = posemb.permute(0, 2, 3, 1).reshape(1, -1, embed_dim)
    posemb = posemb.to(orig_dtype)

    if posemb_prefix is not None:
        posemb = torch.cat([posemb_prefix, posemb], dim=1)
    return posemb
