"""TEST INFRASTRUCTURE ONLY — CPU fp32 restatement of the reference's Depth Pro hot path.

This file is the checker for the CUDA engine.  Only ``tests/``, ``__graft_entry__.smoke()``
and the ``cpu_baseline`` / ``--impl reference`` legs of ``bench.py`` may import it; the
product path (``ml-depth-pro-video_b200/``) never does.

It restates, as plain functions over a reference-format ``state_dict`` (no nn.Module), the
algorithm of ``/root/reference/src/depth_pro`` plus the third-party timm ViT it calls
(un-vendored, unpinned: ``pyproject.toml:9``; restated from timm's published
``vision_transformer.py`` — see ``oracle/timm/__init__.py``).  Every function cites the
reference lines it follows.  Unlike the reference it travels to the GPU box
(``/root/reference`` does not exist there).

Pinning: the reference ships no tests, golden vectors or KATs for this path (SURVEY.md §4,
§8c) — *parity is unpinned by the reference's own tests*.  This restatement is pinned
instead against the reference ITSELF, executed in the build container through
``oracle/reference_loader.py`` (``tests/test_oracle.py::test_oracle_vs_live_reference``), and against the
fixtures that run produced (``tests/golden/*.npz``, made by ``oracle/make_golden.py``).
"""

from __future__ import annotations

import math
from typing import Dict, List, Optional, Tuple

import numpy as np
import torch
import torch.nn.functional as F

IMG = 1536
PATCH = 384
GRID = 24
EMBED = 1024
HEADS = 16


# --------------------------------------------------------------------------------------
# Synthetic inputs (BASELINE.md §3 / SURVEY.md §8d) — shared by tests, golden, bench.
# --------------------------------------------------------------------------------------
def synthetic_image_1536(seed: int = 1) -> torch.Tensor:
    """Config 1: rand(3,1536,1536)*2-1 blended 50/50 with a smooth field; float32 CHW in [-1,1]."""
    g = torch.Generator().manual_seed(seed)
    noise = torch.rand(3, IMG, IMG, generator=g) * 2 - 1
    lin = torch.linspace(-1, 1, IMG)
    v, u = torch.meshgrid(lin, lin, indexing="ij")
    smooth = torch.stack([torch.sin(3 * u + v), torch.cos(2 * v - u), u * v])
    return (0.5 * noise + 0.5 * smooth).contiguous()


def synthetic_frame_u8(index: int, height: int = 1080, width: int = 1920, seed: int = 7) -> np.ndarray:
    """Config 3/4: uint8 HWC frame = moving low-frequency gradient + N(0,8) noise, clipped."""
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


def transform_u8(img_hwc_u8: np.ndarray) -> torch.Tensor:
    """`ToTensor -> Normalize(.5,.5)` (depth_pro.py:125-132): u8 HWC -> f32 CHW in [-1,1]."""
    x = torch.from_numpy(np.ascontiguousarray(img_hwc_u8)).permute(2, 0, 1).to(torch.float32).div(255)
    return (x - 0.5) / 0.5


# --------------------------------------------------------------------------------------
# Encoder plumbing (integer indexing — bit-exact)
# --------------------------------------------------------------------------------------
def create_pyramid(x: torch.Tensor) -> Tuple[torch.Tensor, torch.Tensor, torch.Tensor]:
    """encoder.py:151-168."""
    x1 = F.interpolate(x, size=None, scale_factor=0.5, mode="bilinear", align_corners=False)
    x2 = F.interpolate(x, size=None, scale_factor=0.25, mode="bilinear", align_corners=False)
    return x, x1, x2


def split(x: torch.Tensor, overlap_ratio: float = 0.25) -> torch.Tensor:
    """encoder.py:170-188 — sliding 384 windows, row-major (j outer), concatenated on dim 0."""
    patch_stride = int(PATCH * (1 - overlap_ratio))
    image_size = x.shape[-1]
    steps = int(math.ceil((image_size - PATCH) / patch_stride)) + 1
    out = []
    for j in range(steps):
        j0 = j * patch_stride
        for i in range(steps):
            i0 = i * patch_stride
            out.append(x[..., j0:j0 + PATCH, i0:i0 + PATCH])
    return torch.cat(out, dim=0)


def merge(x: torch.Tensor, batch_size: int, padding: int = 3) -> torch.Tensor:
    """encoder.py:190-217 — crop `padding` off every interior edge and stitch."""
    steps = int(math.sqrt(x.shape[0] // batch_size))
    idx = 0
    rows = []
    for j in range(steps):
        row = []
        for i in range(steps):
            o = x[batch_size * idx: batch_size * (idx + 1)]
            if j != 0:
                o = o[..., padding:, :]
            if i != 0:
                o = o[..., :, padding:]
            if j != steps - 1:
                o = o[..., :-padding, :]
            if i != steps - 1:
                o = o[..., :, :-padding]
            row.append(o)
            idx += 1
        rows.append(torch.cat(row, dim=-1))
    return torch.cat(rows, dim=-2)


def reshape_feature(emb: torch.Tensor, width: int = GRID, height: int = GRID) -> torch.Tensor:
    """encoder.py:219-231 — drop cls, (b, hw, c) -> (b, c, h, w)."""
    b, _, c = emb.shape
    return emb[:, 1:, :].reshape(b, height, width, c).permute(0, 3, 1, 2)


# --------------------------------------------------------------------------------------
# timm ViT-L/16 forward_features (third-party, restated)
# --------------------------------------------------------------------------------------
def vit_forward(sd: Dict[str, torch.Tensor], prefix: str, x: torch.Tensor,
                hook_ids: Tuple[int, ...] = ()) -> Tuple[torch.Tensor, List[torch.Tensor]]:
    """timm VisionTransformer.forward_features (wired at vit_factory.py:97-110, vit.py:33).

    Returns (normed tokens (n,577,1024), [block outputs at hook_ids] — pre final norm,
    as captured by the forward hooks at encoder.py:133-144).
    """
    w = lambda k: sd[prefix + k]
    n = x.shape[0]
    t = F.conv2d(x, w("patch_embed.proj.weight"), w("patch_embed.proj.bias"), stride=16)
    t = t.permute(0, 2, 3, 1).reshape(n, GRID * GRID, EMBED)
    t = torch.cat([w("cls_token").expand(n, -1, -1), t], dim=1) + w("pos_embed")
    hooks = []
    for i in range(24):
        b = f"blocks.{i}."
        h = F.layer_norm(t, (EMBED,), w(b + "norm1.weight"), w(b + "norm1.bias"), 1e-6)
        qkv = F.linear(h, w(b + "attn.qkv.weight"), w(b + "attn.qkv.bias"))
        qkv = qkv.reshape(n, -1, 3, HEADS, EMBED // HEADS).permute(2, 0, 3, 1, 4)
        a = F.scaled_dot_product_attention(qkv[0], qkv[1], qkv[2])
        a = a.transpose(1, 2).reshape(n, -1, EMBED)
        a = F.linear(a, w(b + "attn.proj.weight"), w(b + "attn.proj.bias"))
        t = t + a * w(b + "ls1.gamma")
        h = F.layer_norm(t, (EMBED,), w(b + "norm2.weight"), w(b + "norm2.bias"), 1e-6)
        h = F.gelu(F.linear(h, w(b + "mlp.fc1.weight"), w(b + "mlp.fc1.bias")))
        h = F.linear(h, w(b + "mlp.fc2.weight"), w(b + "mlp.fc2.bias"))
        t = t + h * w(b + "ls2.gamma")
        if i in hook_ids:
            hooks.append(t)
    t = F.layer_norm(t, (EMBED,), w("norm.weight"), w("norm.bias"), 1e-6)
    return t, hooks


# --------------------------------------------------------------------------------------
# DepthProEncoder / MultiresConvDecoder / head / FOV
# --------------------------------------------------------------------------------------
def _project_upsample(sd, name: str, x: torch.Tensor, n_up: int) -> torch.Tensor:
    """encoder.py:60-91 — 1x1 conv (no bias) then n_up ConvTranspose2d k2 s2 (no bias)."""
    x = F.conv2d(x, sd[f"encoder.{name}.0.weight"])
    for i in range(1, n_up + 1):
        x = F.conv_transpose2d(x, sd[f"encoder.{name}.{i}.weight"], stride=2)
    return x


def encoder_forward(sd, x: torch.Tensor, taps: Optional[dict] = None) -> List[torch.Tensor]:
    """encoder.py:233-332."""
    B = x.shape[0]
    x0, x1, x2 = create_pyramid(x)
    x0p = split(x0, 0.25)
    x1p = split(x1, 0.5)
    patches = torch.cat((x0p, x1p, x2), dim=0)
    enc, (hook0, hook1) = vit_forward(sd, "encoder.patch_encoder.", patches, hook_ids=(5, 11))
    enc = reshape_feature(enc)
    lat0 = merge(reshape_feature(hook0)[: B * 25], B, 3)
    lat1 = merge(reshape_feature(hook1)[: B * 25], B, 3)
    e0, e1, e2 = torch.split(enc, [len(x0p), len(x1p), len(x2)], dim=0)
    f0 = merge(e0, B, 3)
    f1 = merge(e1, B, 6)
    f2 = e2
    glob, _ = vit_forward(sd, "encoder.image_encoder.", x2)
    glob = reshape_feature(glob)
    if taps is not None:
        taps.update(patches=patches, lat0_merged=lat0, lat1_merged=lat1, x0_merged=f0,
                    x1_merged=f1, x2_tokens=f2, global_tokens=glob)
    lat0 = _project_upsample(sd, "upsample_latent0", lat0, 3)
    lat1 = _project_upsample(sd, "upsample_latent1", lat1, 2)
    f0 = _project_upsample(sd, "upsample0", f0, 1)
    f1 = _project_upsample(sd, "upsample1", f1, 1)
    f2 = _project_upsample(sd, "upsample2", f2, 1)
    glob = F.conv_transpose2d(glob, sd["encoder.upsample_lowres.weight"],
                              sd["encoder.upsample_lowres.bias"], stride=2)
    glob = F.conv2d(torch.cat((f2, glob), dim=1), sd["encoder.fuse_lowres.weight"],
                    sd["encoder.fuse_lowres.bias"])
    return [lat0, lat1, f0, f1, glob]


def _residual_block(sd, p: str, x: torch.Tensor) -> torch.Tensor:
    """decoder.py:96-118, 182-206 — x + conv(relu(conv(relu(x)))) with biases."""
    d = F.conv2d(F.relu(x), sd[p + "residual.1.weight"], sd[p + "residual.1.bias"], padding=1)
    d = F.conv2d(F.relu(d), sd[p + "residual.3.weight"], sd[p + "residual.3.bias"], padding=1)
    return x + d


def _fusion(sd, f: int, x0: torch.Tensor, x1: Optional[torch.Tensor]) -> torch.Tensor:
    """decoder.py:166-180."""
    p = f"decoder.fusions.{f}."
    x = x0
    if x1 is not None:
        x = x + _residual_block(sd, p + "resnet1.", x1)
    x = _residual_block(sd, p + "resnet2.", x)
    if f != 0:
        x = F.conv_transpose2d(x, sd[p + "deconv.weight"], stride=2)
    return F.conv2d(x, sd[p + "out_conv.weight"], sd[p + "out_conv.bias"])


def decoder_forward(sd, enc: List[torch.Tensor]) -> Tuple[torch.Tensor, torch.Tensor]:
    """decoder.py:74-93."""
    feat = F.conv2d(enc[4], sd["decoder.convs.4.weight"], padding=1)
    lowres = feat
    feat = _fusion(sd, 4, feat, None)
    for i in range(3, -1, -1):
        fi = enc[i] if i == 0 else F.conv2d(enc[i], sd[f"decoder.convs.{i}.weight"], padding=1)
        feat = _fusion(sd, i, feat, fi)
    return feat, lowres


def head_forward(sd, feat: torch.Tensor) -> torch.Tensor:
    """depth_pro.py:182-204."""
    x = F.conv2d(feat, sd["head.0.weight"], sd["head.0.bias"], padding=1)
    x = F.conv_transpose2d(x, sd["head.1.weight"], sd["head.1.bias"], stride=2)
    x = F.relu(F.conv2d(x, sd["head.2.weight"], sd["head.2.bias"], padding=1))
    return F.relu(F.conv2d(x, sd["head.4.weight"], sd["head.4.bias"]))


def fov_forward(sd, x: torch.Tensor, lowres: torch.Tensor) -> Optional[torch.Tensor]:
    """fov.py:56-82.  Three configurations, told apart by the state_dict keys like the reference's `hasattr` checks:
    the default (own ViT encoder, fov.py:47-54), the head without encoder (`fov_encoder_preset=None`: `head =
    fov_head0 + fov_head` applied to the low-resolution feature, fov.py:55-56, 80-82) and no FOV network at all
    (`use_fov_head=False`, depth_pro.py:236-239: `fov_deg` stays None)."""
    if not any(k.startswith("fov.") for k in sd):
        return None
    if "fov.encoder.0.cls_token" not in sd:
        h = lowres
        for i in (0, 2, 4):
            h = F.relu(F.conv2d(h, sd[f"fov.head.{i}.weight"], sd[f"fov.head.{i}.bias"], stride=2, padding=1))
        return F.conv2d(h, sd["fov.head.6.weight"], sd["fov.head.6.bias"])
    x = F.interpolate(x, size=None, scale_factor=0.25, mode="bilinear", align_corners=False)
    t, _ = vit_forward(sd, "fov.encoder.0.", x)
    t = F.linear(t, sd["fov.encoder.1.weight"], sd["fov.encoder.1.bias"])
    t = t[:, 1:].permute(0, 2, 1)
    low = F.relu(F.conv2d(lowres, sd["fov.downsample.0.weight"], sd["fov.downsample.0.bias"],
                          stride=2, padding=1))
    x = t.reshape_as(low) + low
    x = F.relu(F.conv2d(x, sd["fov.head.0.weight"], sd["fov.head.0.bias"], stride=2, padding=1))
    x = F.relu(F.conv2d(x, sd["fov.head.2.weight"], sd["fov.head.2.bias"], stride=2, padding=1))
    return F.conv2d(x, sd["fov.head.4.weight"], sd["fov.head.4.bias"])


@torch.no_grad()
def forward(sd, x: torch.Tensor, taps: Optional[dict] = None) -> Tuple[torch.Tensor, torch.Tensor]:
    """depth_pro.py:218-241 — (canonical inverse depth (B,1,1536,1536), fov_deg (B,1,1,1))."""
    assert x.shape[-1] == IMG and x.shape[-2] == IMG
    enc = encoder_forward(sd, x, taps)
    feat, lowres = decoder_forward(sd, enc)
    canon = head_forward(sd, feat)
    fov = fov_forward(sd, x, lowres)
    if taps is not None:
        taps.update(enc0=enc[0], enc1=enc[1], enc2=enc[2], enc3=enc[3], enc4=enc[4],
                    decoder_out=feat, lowres=lowres, canonical_inverse_depth=canon, fov_deg=fov)
    return canon, fov


@torch.no_grad()
def infer(sd, x: torch.Tensor, f_px=None, taps: Optional[dict] = None,
          interpolation_mode: str = "bilinear") -> Dict[str, torch.Tensor]:
    """depth_pro.py:243-298."""
    if x.dim() == 3:
        x = x.unsqueeze(0)
    _, _, H, W = x.shape
    resize = H != IMG or W != IMG
    if resize:
        x = F.interpolate(x, size=(IMG, IMG), mode=interpolation_mode, align_corners=False)
    canon, fov = forward(sd, x, taps)
    if f_px is None:
        f_px = 0.5 * W / torch.tan(0.5 * torch.deg2rad(fov.to(torch.float)))
    elif not torch.is_tensor(f_px):
        f_px = torch.as_tensor(float(f_px), dtype=torch.float32)
    inv = canon * (W / f_px)
    f_px = f_px.squeeze()
    if resize:
        inv = F.interpolate(inv, size=(H, W), mode=interpolation_mode, align_corners=False)
    depth = 1.0 / torch.clamp(inv, min=1e-4, max=1e4)
    return {"depth": depth.squeeze(), "focallength_px": f_px}


# --------------------------------------------------------------------------------------
# Video add-on geometry / colourising (numpy, as in the reference scripts)
# --------------------------------------------------------------------------------------
def depth_to_3d(depth_in, focallength_px, width, height):
    """img_to_normalized_pointcloud.py:819-856 — pinhole unprojection, x and y negated,
    cx = W/2, cy = H/2, row-major compaction by `valid`; float64 (N,3)."""
    depth_np = np.asarray(depth_in)
    y_idx, x_idx = np.indices((height, width))
    cx = width / 2
    cy = height / 2
    valid = ~np.isnan(depth_np) & (depth_np > 0)
    z = depth_np[valid].flatten()
    x = -1 * (x_idx[valid] - cx) * z / focallength_px
    y = -1 * (y_idx[valid] - cy) * z / focallength_px
    return np.column_stack((x, y, z)), valid


def normalize_depth(depth: np.ndarray, min_depth=None, max_depth=None) -> np.ndarray:
    """generate_depth_maps.py:28-35 — (d - min) / (max - min), clipped to [0,1]; min / max default to nanmin / nanmax."""
    lo = np.nanmin(depth) if min_depth is None else min_depth
    hi = np.nanmax(depth) if max_depth is None else max_depth
    return np.clip((depth - lo) / (hi - lo), 0, 1)


def depth_to_u16(depth: np.ndarray) -> np.ndarray:
    """generate_depth_maps.py:136-139 — 16-bit normalised raw depth."""
    lo, hi = np.nanmin(depth), np.nanmax(depth)
    return ((depth - lo) / (hi - lo) * 65535).astype(np.uint16)


def synthetic_room_points(n: int = 20000, seed: int = 3, tilt_deg: float = 12.0):
    """Seeded stand-in for an unprojected indoor frame: a noisy floor seen by a tilted camera plus boxes and
    walls above it.  Returns (float32 (n,3) points, unit normal (3,), d) with the floor on normal . p + d = 0."""
    rng = np.random.default_rng(seed)
    k = n // 2
    floor = np.column_stack((rng.uniform(-4, 4, k), rng.normal(0, 0.012, k), rng.uniform(1, 9, k)))
    floor[: k // 8, 1] += rng.uniform(0.02, 0.18, k // 8)          # bumps: cells whose low points float above y = 0
    boxes = np.column_stack((rng.uniform(-3, 3, n - k), rng.uniform(-0.3, 2.4, n - k), rng.uniform(2, 8, n - k)))
    pts = np.vstack((floor, boxes))
    a = np.deg2rad(tilt_deg)
    rot = np.array([[1, 0, 0], [0, np.cos(a), -np.sin(a)], [0, np.sin(a), np.cos(a)]])
    cam = pts @ rot.T + np.array([0.0, -1.4, 0.0])                 # camera 1.4 m above the floor, pitched down
    normal = rot @ np.array([0.0, 1.0, 0.0])
    d = 1.4 * normal[1] - normal @ np.array([0.0, 0.0, 0.0])       # floor point (0,0,0) maps to (0,-1.4,0)
    return cam.astype(np.float32), normal, float(d)


def normalize_point_cloud_to_ground(points_3d, normal, d):
    """img_to_normalized_pointcloud.py:880-975 restated.  Signed distances with the unit normal (:873-878);
    Rodrigues rotation of the normal onto +y unless |normal.y| > 0.99 (:912-934); plane to y = const via the
    rotated RAW normal (:939-944); 2nd percentile of the heights within 0.1 of the plane becomes y = 0 when more
    than 10 such points exist (:947-954); points within 0.05 of the plane and below 0 -> 0, other points below
    -0.1 -> -0.1 (:958-972).  float64 in, float64 out."""
    p = np.asarray(points_3d, dtype=np.float64)
    normal = np.asarray(normal, dtype=np.float64)
    unit = normal / np.linalg.norm(normal)
    dist = p @ unit + d
    up = np.array([0.0, 1.0, 0.0])
    if abs(float(normal @ up)) > 0.99:
        out = p.copy()
    else:
        axis = np.cross(unit, up)
        axis /= np.linalg.norm(axis)
        ang = np.arccos(np.clip(unit @ up, -1.0, 1.0))
        K = np.array([[0, -axis[2], axis[1]], [axis[2], 0, -axis[0]], [-axis[1], axis[0], 0]])
        R = np.eye(3) + np.sin(ang) * K + (1 - np.cos(ang)) * (K @ K)
        out = p @ R.T
        out[:, 1] -= -d / (R @ normal)[1]
    near = out[np.abs(dist) < 0.1, 1]
    if near.size > 10:
        out[:, 1] -= np.percentile(near, 2)
    ground = np.abs(dist) < 0.05
    out[(out[:, 1] < 0) & ground, 1] = 0.0
    out[(out[:, 1] < -0.1) & ~ground, 1] = -0.1
    return out


def grid_based_ground_adjustment(points_3d, grid_size=20, percentile=5):
    """img_to_normalized_pointcloud.py:977-1118 restated with a segmented formulation instead of the per-cell
    loop: cells = digitize on linspace edges over the XZ bounds, clipped (:1009-1031); a cell qualifies with >= 10
    points (:1046) and >= 5 heights below 0.2 (:1058-1062); p = percentile of those low heights (:1068); if
    p > 0.01 every point of the cell is lowered by p (y < 0.1), by p * (1 - (y - 0.1) / 1.4) (0.1 <= y < 1.5) or not
    at all, then clamped at 0 (:1075-1106)."""
    p = np.asarray(points_3d, dtype=np.float64)
    out = p.copy()
    x, y, z = p[:, 0], p[:, 1], p[:, 2]
    xe = np.linspace(x.min(), x.max(), grid_size + 1)
    ze = np.linspace(z.min(), z.max(), grid_size + 1)
    cell = np.clip(np.digitize(x, xe) - 1, 0, grid_size - 1) * grid_size + np.clip(np.digitize(z, ze) - 1, 0, grid_size - 1)
    order = np.argsort(cell, kind="stable")
    bounds = np.flatnonzero(np.diff(cell[order], prepend=-1, append=grid_size * grid_size + 1))
    for a, b in zip(bounds[:-1], bounds[1:]):
        idx = order[a:b]
        if idx.size < 10:
            continue
        cy = y[idx]
        low = cy[cy < 0.2]
        if low.size < 5:
            continue
        pc = np.percentile(low, percentile)
        if pc > 0.01:
            adj = np.where(cy < 0.1, pc, np.where(cy < 1.5, pc * (1.0 - (cy - 0.1) / 1.4), 0.0))
            out[idx, 1] = np.maximum(cy - adj, 0.0)
    return out
