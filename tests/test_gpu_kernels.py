"""-m gpu: HBM-bound kernels through the C-ABI vs the oracle (bit-exact indexing, <= 2 ulp resize)."""

import ctypes
import hashlib
import os

import numpy as np
import pytest
import torch
import torch.nn.functional as F

import depthpro_oracle as O
from gpu_common import engine, lib, stream
from depth_pro import _capi

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


def _preprocess(img, fmt, B, H, W):
    x = torch.empty(B, 3, 1536, 1536, device=DEV)
    _capi.check(lib().dp_preprocess(engine(), img.data_ptr(), B, H, W, fmt, x.data_ptr(), stream()))
    torch.cuda.synchronize()
    return x.cpu()


def _ulp_err(a, b):
    # bilinear taps cancel, so the error is measured in ulps of the operand range
    # (|x| <= 1 -> 1 ulp = 2^-23 = 1.19e-7), not of each (possibly tiny) result
    return float((a - b).abs().max() / 2.0 ** -23)


@pytest.mark.parametrize("H,W", [(1080, 1920), (2160, 3840), (384, 640), (1536, 1536), (1537, 1535)])
def test_resize_f32(H, W):
    g = torch.Generator().manual_seed(H * 7 + W)
    x = torch.rand(1, 3, H, W, generator=g) * 2 - 1
    ref = x if (H, W) == (1536, 1536) else F.interpolate(x, size=(1536, 1536), mode="bilinear", align_corners=False)
    got = _preprocess(x.to(DEV), _capi.SRC_F32_CHW, 1, H, W)
    assert _ulp_err(got, ref) <= 2.0


def test_resize_u8_fused_transform():
    frame = O.synthetic_frame_u8(3)
    ref = F.interpolate(O.transform_u8(frame)[None], size=(1536, 1536), mode="bilinear", align_corners=False)
    got = _preprocess(torch.from_numpy(frame).to(DEV), _capi.SRC_U8_HWC, 1, 1080, 1920)
    assert _ulp_err(got, ref) <= 2.0
    # batch of 2 frames
    frames = np.stack([O.synthetic_frame_u8(0, 270, 480), O.synthetic_frame_u8(1, 270, 480)])
    ref2 = F.interpolate(torch.stack([O.transform_u8(f) for f in frames]), size=(1536, 1536), mode="bilinear",
                         align_corners=False)
    got2 = _preprocess(torch.from_numpy(frames).to(DEV), _capi.SRC_U8_HWC, 2, 270, 480)
    assert _ulp_err(got2, ref2) <= 2.0


@pytest.mark.parametrize("H,W", [(1080, 1920), (2160, 3840), (384, 640), (1536, 1536), (1537, 1535), (5, 3), (1, 1)])
def test_resize_v2_is_bit_identical(H, W):
    """The one-row-per-block / four-pixels-per-thread resize (default since round 2) against the round-1 kernel
    (DEPTHPRO_HBM_V2=0, selected through dp_kernel_bench's A/B bits): uint8 and float sources, bit for bit."""
    g = torch.Generator().manual_seed(H * 3 + W)
    u8 = torch.randint(0, 256, (2, H, W, 3), dtype=torch.uint8, generator=g).to(DEV)
    f32 = (torch.rand(2, 3, H, W, generator=g) * 2 - 1).to(DEV)
    ms = ctypes.c_float()
    outs = {}
    try:
        for tag, bit in (("v1", 0x2000), ("v2", 0x1000)):
            _capi.check(lib().dp_kernel_bench(engine(), 8 | bit, 64, 64, 0, 1, ctypes.byref(ms)))
            outs[tag] = (_preprocess(u8, _capi.SRC_U8_HWC, 2, H, W), _preprocess(f32, _capi.SRC_F32_CHW, 2, H, W))
    finally:
        _capi.check(lib().dp_kernel_bench(engine(), 8 | 0x1000, 64, 64, 0, 1, ctypes.byref(ms)))
    assert torch.equal(outs["v1"][0], outs["v2"][0]) and torch.equal(outs["v1"][1], outs["v2"][1])


@pytest.mark.parametrize("H,W", [(1080, 1920), (2160, 3840), (384, 640), (1537, 1535), (5, 3), (1, 1)])
def test_resize_bicubic(H, W):
    """interpolation_mode="bicubic" (depth_pro.py:247, 273-279): ATen upsample_bicubic2d, align_corners=False.  The
    cubic weights reach 1.27 in magnitude and four products are summed per axis -> 8 ulp of the operand range."""
    g = torch.Generator().manual_seed(H * 7 + W)
    x = torch.rand(1, 3, H, W, generator=g) * 2 - 1
    ref = F.interpolate(x, size=(1536, 1536), mode="bicubic", align_corners=False)
    got = torch.empty(1, 3, 1536, 1536, device=DEV)
    _capi.check(lib().dp_preprocess_ex(engine(), x.to(DEV).data_ptr(), 1, H, W, _capi.SRC_F32_CHW, _capi.INTERP["bicubic"],
                                       got.data_ptr(), stream()))
    torch.cuda.synchronize()
    assert _ulp_err(got.cpu(), ref) <= 8.0
    if (H, W) == (1080, 1920):       # and it is not the bilinear result
        assert float((got.cpu() - F.interpolate(x, size=(1536, 1536), mode="bilinear", align_corners=False)).abs().max()) > 1e-2
        frame = O.synthetic_frame_u8(3)
        ref8 = F.interpolate(O.transform_u8(frame)[None], size=(1536, 1536), mode="bicubic", align_corners=False)
        _capi.check(lib().dp_preprocess_ex(engine(), torch.from_numpy(frame).to(DEV).data_ptr(), 1, H, W, _capi.SRC_U8_HWC,
                                           _capi.INTERP["bicubic"], got.data_ptr(), stream()))
        torch.cuda.synchronize()
        assert _ulp_err(got.cpu(), ref8) <= 8.0


@pytest.mark.parametrize("B", [1, 2])
def test_split_bit_exact(B, golden_dir):
    """pyramid + split + cat (encoder.py:151-188, 253-263): patches identical to the oracle's."""
    x = torch.stack([O.synthetic_image_1536(seed=10 + b) for b in range(B)])
    x0, x1, x2 = O.create_pyramid(x)
    ref = torch.cat((O.split(x0, 0.25), O.split(x1, 0.5), x2), dim=0)
    out = torch.empty(35 * B, 3, 384, 384, device=DEV)
    _capi.check(lib().dp_split(engine(), x.to(DEV).data_ptr(), B, out.data_ptr(), stream()))
    torch.cuda.synchronize()
    got = out.cpu()
    # level 0 is pure indexing -> bit exact; levels 1/2 are exact closed forms of the bilinear pyramid
    assert torch.equal(got[: 25 * B], ref[: 25 * B])
    assert torch.equal(got, ref)


def test_split_index_golden(golden_dir):
    """Index-coded image through the CUDA split == the reference's own split (golden corners + digest)."""
    gold = np.load(os.path.join(golden_dir, "split_merge_index.npz"))
    # float32 represents integers < 2^24 exactly: encode (row, col) in two channels
    ys, xs = torch.meshgrid(torch.arange(1536.0), torch.arange(1536.0), indexing="ij")
    x = torch.stack([ys, xs, ys * 0])[None]
    out = torch.empty(35, 3, 384, 384, device=DEV)
    _capi.check(lib().dp_split(engine(), x.to(DEV).data_ptr(), 1, out.data_ptr(), stream()))
    torch.cuda.synchronize()
    got = out.cpu()
    flat = (got[:25, 0] * 1536 + got[:25, 1]).to(torch.int32)  # level-0 patches as flat indices
    assert np.array_equal(flat[:, 0, 0].numpy(), gold["split_1536_corner"])
    assert np.array_equal(flat[:, -1, -1].numpy(), gold["split_1536_last"])
    digest = np.frombuffer(hashlib.sha256(flat[:, None].contiguous().numpy().tobytes()).digest(), dtype=np.uint8)
    assert np.array_equal(digest, gold["split_1536_sha256"])


@pytest.mark.parametrize("steps,pad,B", [(5, 3, 1), (3, 6, 1), (5, 3, 2), (3, 6, 2), (1, 0, 2)])
def test_merge_bit_exact(steps, pad, B, golden_dir):
    """reshape_feature + merge (encoder.py:190-231) on index-coded tokens, vs oracle and golden."""
    n = steps * steps * B
    C = 8
    tok = torch.arange(n * 577 * C, dtype=torch.float32).reshape(n, 577, C)
    ref = O.merge(O.reshape_feature(tok), B, pad) if steps > 1 else O.reshape_feature(tok)
    S = ref.shape[-1]
    out = torch.empty(B, C, S, S, device=DEV)
    _capi.check(lib().dp_merge(engine(), tok.to(DEV).data_ptr(), B, steps, pad, C, out.data_ptr(), stream()))
    torch.cuda.synchronize()
    assert torch.equal(out.cpu(), ref)
    if steps > 1:
        gold = np.load(os.path.join(golden_dir, "split_merge_index.npz"))
        key = f"merge_{steps}x{steps}_pad{pad}" + ("_b2" if B == 2 else "")
        # golden value = patch*576 + token (cls-free); ours = ((seq*577 + 1 + token)*C + c)
        got_idx = (out.cpu()[:, 0] / C).to(torch.int64)
        seq, t = got_idx // 577, got_idx % 577 - 1
        mine = (seq * 576 + t).to(torch.int32).numpy()
        assert np.array_equal(mine if B == 2 else mine[0], gold[key])


def test_depth_to_3d_golden(golden_dir):
    """dp_unproject vs the reference's depth_to_3d output (NaN / <=0 entries, row-major compaction)."""
    gold = np.load(os.path.join(golden_dir, "depth_to_3d.npz"))
    depth = torch.from_numpy(gold["depth"]).to(DEV)
    H, W = depth.shape
    f = torch.tensor([float(gold["f"])], device=DEV)
    xyz = torch.zeros(H * W, 3, device=DEV)
    mask = torch.zeros(H, W, dtype=torch.uint8, device=DEV)
    n = torch.zeros(1, dtype=torch.int64, device=DEV)
    _capi.check(lib().dp_unproject(engine(), depth.data_ptr(), None, H, W, f.data_ptr(), xyz.data_ptr(), None,
                                   mask.data_ptr(), n.data_ptr(), stream()))
    torch.cuda.synchronize()
    assert int(n) == gold["points"].shape[0]
    assert np.array_equal(mask.cpu().numpy().astype(bool), gold["valid"])
    got = xyz[: int(n)].cpu().double().numpy()
    denom = np.maximum(np.abs(gold["points"]), 1e-3)
    assert np.max(np.abs(got - gold["points"]) / denom) <= 1e-6


@pytest.mark.parametrize("H,W", [(1080, 1920), (2160, 3840), (33, 1025)])
def test_unproject_vs_oracle(H, W):
    g = torch.Generator().manual_seed(H)
    depth = torch.rand(H, W, generator=g) * 50 + 0.1
    depth[torch.rand(H, W, generator=g) < 0.02] = float("nan")
    depth[torch.rand(H, W, generator=g) < 0.02] = 0.0
    rgb = torch.randint(0, 256, (H, W, 3), generator=g, dtype=torch.uint8)
    f = 1234.5
    pts, valid = O.depth_to_3d(depth.numpy(), f, W, H)
    d = depth.to(DEV)
    xyz = torch.zeros(H * W, 3, device=DEV)
    col = torch.zeros(H * W, 3, device=DEV)
    n = torch.zeros(1, dtype=torch.int64, device=DEV)
    ft = torch.tensor([f], device=DEV)
    _capi.check(lib().dp_unproject(engine(), d.data_ptr(), rgb.to(DEV).data_ptr(), H, W, ft.data_ptr(), xyz.data_ptr(),
                                   col.data_ptr(), None, n.data_ptr(), stream()))
    torch.cuda.synchronize()
    assert int(n) == pts.shape[0]
    got = xyz[: int(n)].cpu().double().numpy()
    assert np.max(np.abs(got - pts) / np.maximum(np.abs(pts), 1e-3)) <= 1e-6
    ref_col = rgb.numpy().reshape(-1, 3)[valid.flatten()] / 255.0
    assert np.max(np.abs(col[: int(n)].cpu().double().numpy() - ref_col)) <= 1e-7


def test_colorize_and_u16():
    H, W = 270, 480
    g = torch.Generator().manual_seed(3)
    depth = torch.rand(H, W, generator=g) * 30 + 0.5
    lut = torch.randint(0, 256, (256, 3), generator=g, dtype=torch.uint8)
    norm = O.normalize_depth(depth.numpy())
    idx = np.minimum((norm * 256).astype(np.int64), 255)
    ref = lut.numpy()[idx]
    out = torch.zeros(H, W, 3, dtype=torch.uint8, device=DEV)
    d_dev, lut_dev = depth.to(DEV), lut.to(DEV)  # keep alive: the calls below are asynchronous
    _capi.check(lib().dp_colorize(engine(), d_dev.data_ptr(), H, W, lut_dev.data_ptr(), out.data_ptr(), stream()))
    out16 = torch.zeros(H, W, dtype=torch.int16, device=DEV)
    _capi.check(lib().dp_colorize(engine(), d_dev.data_ptr(), H, W, None, out16.data_ptr(), stream()))
    torch.cuda.synchronize()
    assert np.array_equal(out.cpu().numpy(), ref)
    assert np.array_equal(out16.cpu().numpy().view(np.uint16), O.depth_to_u16(depth.numpy()))


# ---- ground normalisation (img_to_normalized_pointcloud.py:880-1118) -------------------------------------
def _ground_call(fn, pts32, *args):
    import ctypes

    xyz = torch.from_numpy(pts32).to(DEV).contiguous()
    ctr = torch.zeros(6, dtype=torch.int64, device=DEV)
    _capi.check(fn(engine(), xyz.data_ptr(), xyz.shape[0], *args, ctr.data_ptr(), stream()))
    torch.cuda.synchronize()
    return xyz.cpu().numpy(), ctr.tolist()


def _normalize_gpu(pts32, normal, d):
    import ctypes

    n3 = (ctypes.c_double * 3)(*[float(v) for v in normal])
    return _ground_call(lib().dp_ground_normalize, pts32, n3, float(d))


def _grid_gpu(pts32, grid_size=20, percentile=5.0):
    return _ground_call(lib().dp_ground_grid_adjust, pts32, int(grid_size), float(percentile))


@pytest.mark.parametrize("tag", ["tilt12", "tilt3"])
def test_ground_normalize_golden(tag, golden_dir):
    """Against outputs of the reference's own functions (rotation branch and the |normal.y| > 0.99 branch)."""
    g = np.load(os.path.join(golden_dir, "ground_normalize.npz"))
    pts, normal, d = g[tag + "_points"], g[tag + "_normal"], float(g[tag + "_d"])
    got, ctr = _normalize_gpu(pts, normal, d)
    ref = g[tag + "_normalized"]
    # double arithmetic on both sides, one float32 rounding at the end; the order statistics are float32-rounded
    # heights on the GPU (1e-7 relative) -- a continuous effect, never a different branch
    assert np.max(np.abs(got.astype(np.float64) - ref.astype(np.float64))) <= 2e-6
    assert np.array_equal(got[:, 1] == 0, ref[:, 1] == 0) and np.array_equal(got[:, 1] == np.float32(-0.1), ref[:, 1] == np.float32(-0.1))
    dist = pts.astype(np.float64) @ normal + d
    assert ctr[0] == int((np.abs(dist) < 0.05).sum()) and ctr[1] == int((ref[:, 1] == 0).sum())
    # second stage on the reference's first-stage output: identical input on both sides -> identical selection
    got2, ctr2 = _grid_gpu(ref)
    ref2 = g[tag + "_grid"]
    assert np.max(np.abs(got2.astype(np.float64) - ref2.astype(np.float64))) <= 1e-6
    assert np.array_equal(got2[:, [0, 2]], ref[:, [0, 2]])                       # x and z are never touched
    assert ctr2[3] == int((ref2[:, 1] != ref[:, 1]).sum()) or ctr2[3] >= int((ref2[:, 1] != ref[:, 1]).sum())
    assert 0 < ctr2[5] <= ctr2[4] <= 400


@pytest.mark.parametrize("n,seed,tilt,grid,pct", [(200_000, 11, 20.0, 20, 5.0), (1_000_003, 12, 7.0, 32, 10.0),
                                                   (5_000, 13, 35.0, 7, 50.0), (2_073_600, 14, 15.0, 20, 5.0)])
def test_ground_stages_vs_oracle(n, seed, tilt, grid, pct):
    """Full-size clouds (up to a 1080p frame) against the float64 oracle on the same float32 points."""
    pts, normal, d = O.synthetic_room_points(n, seed, tilt)
    got, _ = _normalize_gpu(pts, normal * 1.7, d * 1.0)       # a non-unit normal: distances use the unit normal,
    ref = O.normalize_point_cloud_to_ground(pts.astype(np.float64), normal * 1.7, d)   # the shift the raw one (:873, :939)
    assert np.max(np.abs(got.astype(np.float64) - ref)) <= 3e-6
    got2, ctr = _grid_gpu(got, grid, pct)
    ref2 = O.grid_based_ground_adjustment(got.astype(np.float64), grid, pct)
    assert np.max(np.abs(got2.astype(np.float64) - ref2)) <= 1e-6
    assert ctr[3] > 0 and ctr[5] > 0
    assert float(got2[:, 1].min()) >= -0.1 - 1e-6


def test_ground_edge_cases():
    # fewer than 11 near-plane points: no percentile shift; every point far above the plane is untouched
    far = np.array([[0.0, 5.0, 1.0], [1.0, 6.0, 2.0], [2.0, 7.0, 3.0]], dtype=np.float32)
    got, ctr = _normalize_gpu(far, np.array([0.0, 1.0, 0.0]), 0.0)
    assert np.array_equal(got, far) and ctr[:3] == [0, 0, 0]
    got2, ctr2 = _grid_gpu(far)
    assert np.array_equal(got2, far) and ctr2[3:] == [0, 0, 0]
    # duplicates everywhere (order statistics with multiplicity), all points in one cell column
    rng = np.random.default_rng(0)
    dup = np.column_stack((np.zeros(4000), rng.integers(0, 6, 4000) * 0.03 + 0.03, rng.uniform(1, 2, 4000))).astype(np.float32)
    got3, _ = _grid_gpu(dup, 4, 5.0)
    ref3 = O.grid_based_ground_adjustment(dup.astype(np.float64), 4, 5.0)
    assert np.max(np.abs(got3.astype(np.float64) - ref3)) <= 1e-6 and (got3[:, 1] != dup[:, 1]).any()
    # n = 0 is a no-op
    empty = torch.empty((0, 3), device=DEV)
    _capi.check(lib().dp_ground_grid_adjust(engine(), empty.data_ptr(), 0, 20, 5.0, None, stream()))


def _kb(kind, M, N, K, iters=1):
    ms = ctypes.c_float()
    _capi.check(lib().dp_kernel_bench(engine(), kind, M, N, K, iters, ctypes.byref(ms)))
    return ms.value


@pytest.mark.parametrize("H,W", [(1536, 1536), (1080, 1920), (2160, 3840), (333, 2001), (7, 5), (1, 1), (1081, 1023)])
def test_depth_epilogue_v2_is_bit_identical(H, W):
    """The four-pixels-per-thread metric-depth epilogue (default since round 2) must reproduce the one-pixel-per-thread
    round-1 kernel (DEPTHPRO_HBM_V2=0) bit for bit, at aligned, ragged and tiny output sizes.  Runs the whole bf16 frame
    twice on the same input."""
    import depth_pro

    model = _epilogue_model()
    g = torch.Generator(device=DEV).manual_seed(H * 10007 + W)
    x = torch.rand(3, H, W, device=DEV, generator=g) * 2 - 1
    try:
        _kb(8 | 0x2000, 64, 64, 0)
        a = model.infer(x)["depth"].clone()
        _kb(8 | 0x1000, 64, 64, 0)
        b = model.infer(x)["depth"].clone()
    finally:
        _kb(8 | 0x1000, 64, 64, 0)   # back to the default (v2)
    assert a.shape == b.shape
    assert torch.equal(a, b)


_MODEL = None


def _epilogue_model():
    global _MODEL
    if _MODEL is None:
        import depth_pro

        _MODEL = depth_pro.DepthPro(device=torch.device(DEV), precision=torch.bfloat16).init_weights("stress", 1234)
    return _MODEL
