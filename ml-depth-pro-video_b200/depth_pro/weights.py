"""Parameter manifest and seeded initialisation recipes for Depth Pro.

The engine ingests a plain ``state_dict`` whose keys and shapes equal those of the
reference model (``/root/reference/src/depth_pro/depth_pro.py:72-151``) so
that a real ``depth_pro.pt`` loads unchanged.  ``manifest()`` lists every tensor
(name -> shape) in the reference's registration order; ``tests/golden/
state_dict_manifest.json`` (dumped from the reference itself) pins it.

Checkpoints are not available offline, so two seeded recipes are provided:

* ``reference_like_init`` — the distributions the reference would draw (timm trunc-normal
  0.02, LayerScale 1e-5, PyTorch conv defaults, ``head.4.bias = 0``).  Degenerate: the
  canonical inverse depth is identically 0 (SURVEY.md §7 hard-part 1).
* ``stress_init`` ("recipe B") — variance-preserving weights that make every
  layer contribute to the output; this is what the parity tests and ``bench.py`` use.

Each tensor is drawn from its own CPU generator seeded by ``(seed, crc32(name))`` so the
result is independent of generation order and identical on every host.
"""

from __future__ import annotations

import math
import zlib
from collections import OrderedDict
from typing import Dict, Iterator, Tuple

import torch

EMBED = 1024
DEPTH = 24
MLP = 4096
TOKENS = 577
DEC = 256
VIT_PREFIXES = ("encoder.patch_encoder.", "encoder.image_encoder.", "fov.encoder.0.")


def _vit(prefix: str) -> Iterator[Tuple[str, Tuple[int, ...]]]:
    yield prefix + "cls_token", (1, 1, EMBED)
    yield prefix + "pos_embed", (1, TOKENS, EMBED)
    yield prefix + "patch_embed.proj.weight", (EMBED, 3, 16, 16)
    yield prefix + "patch_embed.proj.bias", (EMBED,)
    for i in range(DEPTH):
        b = f"{prefix}blocks.{i}."
        yield b + "norm1.weight", (EMBED,)
        yield b + "norm1.bias", (EMBED,)
        yield b + "attn.qkv.weight", (3 * EMBED, EMBED)
        yield b + "attn.qkv.bias", (3 * EMBED,)
        yield b + "attn.proj.weight", (EMBED, EMBED)
        yield b + "attn.proj.bias", (EMBED,)
        yield b + "ls1.gamma", (EMBED,)
        yield b + "norm2.weight", (EMBED,)
        yield b + "norm2.bias", (EMBED,)
        yield b + "mlp.fc1.weight", (MLP, EMBED)
        yield b + "mlp.fc1.bias", (MLP,)
        yield b + "mlp.fc2.weight", (EMBED, MLP)
        yield b + "mlp.fc2.bias", (EMBED,)
        yield b + "ls2.gamma", (EMBED,)
    yield prefix + "norm.weight", (EMBED,)
    yield prefix + "norm.bias", (EMBED,)


def manifest(fov: str = "encoder") -> "OrderedDict[str, Tuple[int, ...]]":
    """name -> shape in the reference's state_dict order: all 1119 tensors of the default configuration
    (``fov="encoder"``), or of ``fov="head"`` (``fov_encoder_preset=None``: FOV head without its own ViT, fov.py:55-56)
    / ``fov=None`` (``use_fov_head=False``: no ``fov.*`` tensors, depth_pro.py:100-108)."""
    if fov not in ("encoder", "head", None):
        raise ValueError(f"fov must be 'encoder', 'head' or None, got {fov!r}")
    m: "OrderedDict[str, Tuple[int, ...]]" = OrderedDict()
    for k, s in _vit("encoder.patch_encoder."):
        m[k] = s
    for k, s in _vit("encoder.image_encoder."):
        m[k] = s
    # encoder.py:93-130 — project (1x1, no bias) + ConvTranspose2d k2 s2 (no bias) chains
    m["encoder.upsample_latent0.0.weight"] = (256, EMBED, 1, 1)
    for i in (1, 2, 3):
        m[f"encoder.upsample_latent0.{i}.weight"] = (256, 256, 2, 2)
    m["encoder.upsample_latent1.0.weight"] = (256, EMBED, 1, 1)
    for i in (1, 2):
        m[f"encoder.upsample_latent1.{i}.weight"] = (256, 256, 2, 2)
    for name, d in (("upsample0", 512), ("upsample1", 1024), ("upsample2", 1024)):
        m[f"encoder.{name}.0.weight"] = (d, EMBED, 1, 1)
        m[f"encoder.{name}.1.weight"] = (d, d, 2, 2)
    m["encoder.upsample_lowres.weight"] = (EMBED, 1024, 2, 2)
    m["encoder.upsample_lowres.bias"] = (1024,)
    m["encoder.fuse_lowres.weight"] = (1024, 2048, 1, 1)
    m["encoder.fuse_lowres.bias"] = (1024,)
    # decoder.py:42-72
    for i, d in ((1, 256), (2, 512), (3, 1024), (4, 1024)):
        m[f"decoder.convs.{i}.weight"] = (DEC, d, 3, 3)
    for f in range(5):
        for rn in ("resnet1", "resnet2"):
            for c in (1, 3):
                m[f"decoder.fusions.{f}.{rn}.residual.{c}.weight"] = (DEC, DEC, 3, 3)
                m[f"decoder.fusions.{f}.{rn}.residual.{c}.bias"] = (DEC,)
        if f != 0:
            m[f"decoder.fusions.{f}.deconv.weight"] = (DEC, DEC, 2, 2)
        m[f"decoder.fusions.{f}.out_conv.weight"] = (DEC, DEC, 1, 1)
        m[f"decoder.fusions.{f}.out_conv.bias"] = (DEC,)
    # depth_pro.py:182-204
    m["head.0.weight"] = (128, 256, 3, 3)
    m["head.0.bias"] = (128,)
    m["head.1.weight"] = (128, 128, 2, 2)
    m["head.1.bias"] = (128,)
    m["head.2.weight"] = (32, 128, 3, 3)
    m["head.2.bias"] = (32,)
    m["head.4.weight"] = (1, 32, 1, 1)
    m["head.4.bias"] = (1,)
    # fov.py:29-55
    if fov is None:
        return m
    if fov == "head":
        for i, (co, ci) in zip((0, 2, 4), ((128, 256), (64, 128), (32, 64))):
            m[f"fov.head.{i}.weight"] = (co, ci, 3, 3)
            m[f"fov.head.{i}.bias"] = (co,)
        m["fov.head.6.weight"] = (1, 32, 6, 6)
        m["fov.head.6.bias"] = (1,)
        return m
    for k, s in _vit("fov.encoder.0."):
        m[k] = s
    m["fov.encoder.1.weight"] = (128, EMBED)
    m["fov.encoder.1.bias"] = (128,)
    m["fov.downsample.0.weight"] = (128, 256, 3, 3)
    m["fov.downsample.0.bias"] = (128,)
    m["fov.head.0.weight"] = (64, 128, 3, 3)
    m["fov.head.0.bias"] = (64,)
    m["fov.head.2.weight"] = (32, 64, 3, 3)
    m["fov.head.2.bias"] = (32,)
    m["fov.head.4.weight"] = (1, 32, 6, 6)
    m["fov.head.4.bias"] = (1,)
    return m


def _gen(seed: int, name: str) -> torch.Generator:
    g = torch.Generator(device="cpu")
    g.manual_seed((int(seed) * 1000003 + zlib.crc32(name.encode())) & 0x7FFFFFFFFFFFFFFF)
    return g


def _is_convT(name: str) -> bool:
    if name in ("encoder.upsample_lowres.weight", "head.1.weight"):
        return True
    if name.endswith("deconv.weight"):
        return True
    if name.startswith("encoder.upsample") and name.endswith(".weight"):
        idx = name.split(".")[-2]
        return idx.isdigit() and int(idx) >= 1
    return False


def stress_tensor(name: str, shape: Tuple[int, ...], seed: int) -> torch.Tensor:
    """One tensor of "recipe B" (SURVEY.md §7): fp32, CPU, deterministic in (seed, name)."""
    g = _gen(seed, name)

    def normal(std):
        return torch.empty(shape, dtype=torch.float32).normal_(0.0, std, generator=g)

    def uniform(lo, hi):
        return torch.empty(shape, dtype=torch.float32).uniform_(lo, hi, generator=g)

    leaf = name.rsplit(".", 1)[-1]
    if name == "head.4.bias":
        return torch.full(shape, 2.0)
    if name in ("fov.head.4.bias", "fov.head.6.bias") and shape == (1,):   # the FOV head's last conv (6x6 -> 1)
        return torch.full(shape, 60.0)
    if name == "head.4.weight":
        return normal(0.25 / math.sqrt(32))
    if leaf == "gamma":
        return uniform(0.05, 0.3)
    if leaf in ("cls_token", "pos_embed"):
        return normal(0.02)
    if ".norm" in name and leaf == "weight" and len(shape) == 1:
        return uniform(0.8, 1.2)
    if leaf == "bias":
        return normal(0.02)
    if len(shape) == 2:  # Linear (out, in)
        return normal(1.0 / math.sqrt(shape[1]))
    if len(shape) == 4:
        if _is_convT(name):  # (Cin, Cout, 2, 2): every output pixel sums Cin terms
            return normal(1.0 / math.sqrt(shape[0]))
        fan_in = shape[1] * shape[2] * shape[3]
        return normal(1.0 / math.sqrt(fan_in))
    raise KeyError(f"no init rule for {name} {shape}")


def reference_like_tensor(name: str, shape: Tuple[int, ...], seed: int) -> torch.Tensor:
    """The distributions the reference's own construction draws (degenerate output)."""
    g = _gen(seed, name)
    leaf = name.rsplit(".", 1)[-1]
    t = torch.empty(shape, dtype=torch.float32)
    if leaf == "gamma":
        return t.fill_(1e-5)
    if leaf == "cls_token":
        return t.normal_(0.0, 1e-6, generator=g)
    if leaf == "pos_embed":
        return t.normal_(0.0, 0.02, generator=g).clamp_(-0.04, 0.04)
    if ".norm" in name and len(shape) == 1:
        return t.fill_(1.0 if leaf == "weight" else 0.0)
    if name == "head.4.bias":
        return t.zero_()
    if len(shape) == 2:
        return t.normal_(0.0, 0.02, generator=g).clamp_(-0.04, 0.04)
    is_vit = any(name.startswith(p) for p in VIT_PREFIXES) and "patch_embed" not in name
    if leaf == "bias" and is_vit:
        return t.zero_()
    # PyTorch conv / linear default: U(-1/sqrt(fan_in), 1/sqrt(fan_in))
    # (Conv2d and ConvTranspose2d both use weight.size(1) * kh * kw as fan_in.)
    if len(shape) == 4:
        fan_in = shape[1] * shape[2] * shape[3]
    else:
        fan_in = _bias_fan_in(name)
    b = 1.0 / math.sqrt(fan_in)
    return t.uniform_(-b, b, generator=g)


def _bias_fan_in(name: str) -> int:
    w = name.rsplit(".", 1)[0] + ".weight"
    s = manifest()[w] if w in manifest() else manifest("head")[w]   # fov.head.6 exists in the head-only config only
    if len(s) == 2:
        return s[1]
    return s[1] * s[2] * s[3]


def stress_init(seed: int = 1234, fov: str = "encoder") -> Dict[str, torch.Tensor]:
    """Full "recipe B" state_dict (fp32, CPU, ~3.8 GB)."""
    return OrderedDict((k, stress_tensor(k, s, seed)) for k, s in manifest(fov).items())


def reference_like_init(seed: int = 0, fov: str = "encoder") -> Dict[str, torch.Tensor]:
    return OrderedDict((k, reference_like_tensor(k, s, seed)) for k, s in manifest(fov).items())
