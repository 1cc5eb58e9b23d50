"""TEST INFRASTRUCTURE ONLY — minimal stand-in for the third-party `timm` package.

The reference (`/root/reference/src/depth_pro/network/vit_factory.py:12,97-99`,
`vit.py:5`) depends on `timm` (pyproject.toml:9, unpinned), which is neither vendored in
the reference tree nor installed in this image.  This shim restates the published
algorithm of timm's `VisionTransformer` for the ONE model id the reference uses,
`vit_large_patch14_dinov2` (patch 14, img 518, dim 1024, depth 24, heads 16,
init_values=1e-5, qkv_bias, mlp_ratio 4, exact GELU, LayerNorm eps 1e-6, class token,
no register tokens, num_classes=0), with exactly the attribute surface the reference
touches, so that the reference's own `depth_pro` package imports and runs UNMODIFIED
from `/root/reference/src`.  It is only ever imported by `oracle/reference_loader.py`
(tests / golden-vector generation / the CPU baseline), never by the product path.

Parity status: the timm boundary is pinned by no reference test ("parity unpinned",
SURVEY.md §8c); module / parameter names follow timm so `state_dict()` keys equal those of
a real `depth_pro.pt`.
"""

from __future__ import annotations

import torch
import torch.nn as nn
import torch.nn.functional as F

from .layers import resample_abs_pos_embed

__all__ = ["create_model", "VisionTransformer"]


class PatchEmbed(nn.Module):
    """timm.layers.PatchEmbed with `dynamic_img_size=True` (NHWC output, no size assert).

    `proj` is read at call time because the reference swaps it out
    (`vit.py:95-106`); img_size / patch_size / grid_size are mutable tuples
    (`vit.py:53-56, 109-121`).
    """

    def __init__(self, img_size=518, patch_size=14, in_chans=3, embed_dim=1024):
        super().__init__()
        self.img_size = (img_size, img_size)
        self.patch_size = (patch_size, patch_size)
        self.grid_size = (img_size // patch_size, img_size // patch_size)
        self.num_patches = self.grid_size[0] * self.grid_size[1]
        self.proj = nn.Conv2d(in_chans, embed_dim, kernel_size=patch_size, stride=patch_size)
        self.norm = nn.Identity()

    def forward(self, x):
        x = self.proj(x)
        return x.permute(0, 2, 3, 1)  # NCHW -> NHWC


class Attention(nn.Module):
    def __init__(self, dim, num_heads):
        super().__init__()
        self.num_heads = num_heads
        self.head_dim = dim // num_heads
        self.scale = self.head_dim**-0.5
        self.qkv = nn.Linear(dim, dim * 3, bias=True)
        self.q_norm = nn.Identity()
        self.k_norm = nn.Identity()
        self.attn_drop = nn.Dropout(0.0)
        self.proj = nn.Linear(dim, dim)
        self.proj_drop = nn.Dropout(0.0)

    def forward(self, x):
        B, N, C = x.shape
        qkv = self.qkv(x).reshape(B, N, 3, self.num_heads, self.head_dim).permute(2, 0, 3, 1, 4)
        q, k, v = qkv.unbind(0)
        x = F.scaled_dot_product_attention(q, k, v)
        x = x.transpose(1, 2).reshape(B, N, C)
        return self.proj(x)


class LayerScale(nn.Module):
    def __init__(self, dim, init_values=1e-5):
        super().__init__()
        self.gamma = nn.Parameter(init_values * torch.ones(dim))

    def forward(self, x):
        return x * self.gamma


class Mlp(nn.Module):
    def __init__(self, dim, hidden):
        super().__init__()
        self.fc1 = nn.Linear(dim, hidden)
        self.act = nn.GELU()
        self.drop1 = nn.Dropout(0.0)
        self.norm = nn.Identity()
        self.fc2 = nn.Linear(hidden, dim)
        self.drop2 = nn.Dropout(0.0)

    def forward(self, x):
        return self.fc2(self.act(self.fc1(x)))


class Block(nn.Module):
    def __init__(self, dim, num_heads, mlp_ratio=4.0, init_values=1e-5):
        super().__init__()
        self.norm1 = nn.LayerNorm(dim, eps=1e-6)
        self.attn = Attention(dim, num_heads)
        self.ls1 = LayerScale(dim, init_values)
        self.drop_path1 = nn.Identity()
        self.norm2 = nn.LayerNorm(dim, eps=1e-6)
        self.mlp = Mlp(dim, int(dim * mlp_ratio))
        self.ls2 = LayerScale(dim, init_values)
        self.drop_path2 = nn.Identity()

    def forward(self, x):
        x = x + self.ls1(self.attn(self.norm1(x)))
        x = x + self.ls2(self.mlp(self.norm2(x)))
        return x


class VisionTransformer(nn.Module):
    def __init__(self, img_size=518, patch_size=14, embed_dim=1024, depth=24, num_heads=16,
                 init_values=1e-5, dynamic_img_size=True):
        super().__init__()
        self.num_classes = 0
        self.embed_dim = self.num_features = embed_dim
        self.num_prefix_tokens = 1
        self.num_reg_tokens = 0
        self.has_class_token = True
        self.no_embed_class = False
        self.dynamic_img_size = dynamic_img_size
        self.grad_checkpointing = False

        self.patch_embed = PatchEmbed(img_size, patch_size, 3, embed_dim)
        self.cls_token = nn.Parameter(torch.zeros(1, 1, embed_dim))
        self.pos_embed = nn.Parameter(torch.randn(1, self.patch_embed.num_patches + 1, embed_dim) * 0.02)
        self.pos_drop = nn.Dropout(0.0)
        self.patch_drop = nn.Identity()
        self.norm_pre = nn.Identity()
        self.blocks = nn.Sequential(*[Block(embed_dim, num_heads, 4.0, init_values) for _ in range(depth)])
        self.norm = nn.LayerNorm(embed_dim, eps=1e-6)
        self.fc_norm = nn.Identity()
        self.head_drop = nn.Dropout(0.0)
        self.head = nn.Identity()
        self._init_weights()

    def _init_weights(self):
        nn.init.trunc_normal_(self.pos_embed, std=0.02)
        nn.init.normal_(self.cls_token, std=1e-6)
        for m in self.modules():
            if isinstance(m, nn.Linear):
                nn.init.trunc_normal_(m.weight, std=0.02)
                if m.bias is not None:
                    nn.init.zeros_(m.bias)

    def set_grad_checkpointing(self, enable=True):
        self.grad_checkpointing = enable

    def _pos_embed(self, x):
        if self.dynamic_img_size:
            B, H, W, C = x.shape
            pos_embed = resample_abs_pos_embed(
                self.pos_embed, (H, W),
                num_prefix_tokens=0 if self.no_embed_class else self.num_prefix_tokens)
            x = x.view(B, -1, C)
        else:
            pos_embed = self.pos_embed
        x = torch.cat([self.cls_token.expand(x.shape[0], -1, -1), x], dim=1)
        x = x + pos_embed
        return self.pos_drop(x)

    def forward_features(self, x):
        x = self.patch_embed(x)
        x = self._pos_embed(x)
        x = self.patch_drop(x)
        x = self.norm_pre(x)
        x = self.blocks(x)
        x = self.norm(x)
        return x

    def forward(self, x):
        return self.forward_features(x)


def create_model(model_name, pretrained=False, **kwargs):
    if model_name != "vit_large_patch14_dinov2":
        raise RuntimeError(f"timm shim: unknown model {model_name}")
    if pretrained:
        raise RuntimeError("timm shim: pretrained weights are not available offline")
    return VisionTransformer(dynamic_img_size=kwargs.get("dynamic_img_size", False))
