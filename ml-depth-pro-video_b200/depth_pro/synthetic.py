"""Synthetic inputs of the benchmark configurations.

No datasets or video files are available offline; these seeded generators stand in for them.
(`oracle/depthpro_oracle.py` carries its own copy so the checker stays self-contained;
`tests/test_host.py` asserts the two agree.)
"""

from __future__ import annotations

import numpy as np
import torch

IMG = 1536


def synthetic_image_1536(seed: int = 1) -> torch.Tensor:
    """Config 1: rand(3,1536,1536)*2-1 blended 50/50 with a smooth field; float32 CHW in [-1,1]."""
    g = torch.Generator().manual_seed(seed)
    noise = torch.rand(3, IMG, IMG, generator=g) * 2 - 1
    lin = torch.linspace(-1, 1, IMG)
    v, u = torch.meshgrid(lin, lin, indexing="ij")
    smooth = torch.stack([torch.sin(3 * u + v), torch.cos(2 * v - u), u * v])
    return (0.5 * noise + 0.5 * smooth).contiguous()


def synthetic_frame_u8(index: int, height: int = 1080, width: int = 1920, seed: int = 7) -> np.ndarray:
    """Config 3/4: uint8 HWC video frame = moving low-frequency gradient + N(0,8) noise, clipped."""
    rng = np.random.default_rng(seed * 100003 + index)
    y = np.linspace(0, 1, height, dtype=np.float32)[:, None]
    x = np.linspace(0, 1, width, dtype=np.float32)[None, :]
    ph = 0.05 * index
    base = np.stack([
        127.5 + 100 * np.sin(2 * np.pi * (x + ph)) * np.cos(np.pi * y),
        127.5 + 100 * np.cos(2 * np.pi * (y - ph)) * np.sin(np.pi * x + 0.3),
        255 * (0.5 * x + 0.5 * y) + 0 * ph,
    ], axis=-1).astype(np.float32)
    img = base + rng.normal(0, 8, size=base.shape).astype(np.float32)
    return np.clip(np.rint(img), 0, 255).astype(np.uint8)
