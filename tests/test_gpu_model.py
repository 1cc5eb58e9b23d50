"""-m gpu: the full Depth Pro path through the drop-in API vs the CPU oracle and the golden
fixtures produced by the unmodified reference (tests/golden/reference_outputs.npz).

Tolerances are BASELINE.md §4: fp32 per-pixel depth rel-err <= 1e-4 and f_px <= 1e-4;
bf16 median abs-rel <= 5e-3 (p99 reported), f_px <= 1e-2; multi-frame batches bit-identical
to single frames.
"""

import json
import os

import numpy as np
import pytest
import torch

import depth_pro
import depthpro_oracle as O
from depth_pro import weights

pytestmark = pytest.mark.gpu
DEV = torch.device("cuda:0")
SEED = 1234


@pytest.fixture(scope="module")
def state_dict():
    return weights.stress_init(SEED)


@pytest.fixture(scope="module")
def oracle_run(state_dict):
    torch.set_num_threads(os.cpu_count())
    x = O.synthetic_image_1536(1)
    taps = {}
    out = O.infer(state_dict, x, taps=taps)
    return x, out, taps


def _model(state_dict, precision):
    m = depth_pro.DepthPro(device=DEV, precision=precision)
    m.load_state_dict(state_dict, strict=True)
    return m.eval()


@pytest.fixture(scope="module")
def model_fp32(state_dict):
    return _model(state_dict, torch.float32)


@pytest.fixture(scope="module")
def model_bf16(state_dict):
    return _model(state_dict, torch.bfloat16)


def _pix_rel(a, b):
    return ((a.double() - b.double()).abs() / b.double().abs().clamp_min(1e-30))


def test_oracle_matches_golden(oracle_run, golden_dir):
    """The oracle on THIS host reproduces the reference's outputs recorded in the build container."""
    _, out, taps = oracle_run
    gold = np.load(os.path.join(golden_dir, "reference_outputs.npz"))
    g = torch.from_numpy(gold["depth_1536"])
    # two fp32 CPU runs on different hosts (other core count / oneDNN blocking) differ by a few 1e-5
    assert float(_pix_rel(out["depth"][::16, ::16], g).max()) <= 1e-4
    assert abs(float(out["focallength_px"]) - float(gold["f_px_1536"])) / float(gold["f_px_1536"]) <= 1e-4


def test_fp32_vs_oracle(model_fp32, oracle_run, golden_dir):
    x, ref, taps = oracle_run
    pred = model_fp32.infer(x.to(DEV))
    depth = pred["depth"].cpu()
    rel = _pix_rel(depth, ref["depth"])
    f_rel = abs(float(pred["focallength_px"]) - float(ref["focallength_px"])) / float(ref["focallength_px"])
    print(f"fp32: depth rel-err max {float(rel.max()):.3e} median {float(rel.median()):.3e}; f_px rel {f_rel:.3e}")
    # stage taps first: they localise a failure
    for name in ("lat0_merged", "lat1_merged", "x0_merged", "x1_merged", "x2_tokens", "global_tokens",
                 "enc0", "enc1", "enc2", "enc3", "enc4", "lowres", "decoder_out"):
        r = taps[name]
        got = model_fp32.tap(name).cpu().reshape(r.shape)
        err = float((got - r).abs().max() / r.abs().max())
        print(f"  tap {name:14s} max-abs-err/absmax {err:.3e}")
        assert err <= 1e-4, name
    assert float(rel.max()) <= 1e-4
    assert f_rel <= 1e-4
    assert pred["depth"].shape == (1536, 1536) and pred["focallength_px"].dim() == 0
    # against the reference's own recorded output
    gold = np.load(os.path.join(golden_dir, "reference_outputs.npz"))
    assert float(_pix_rel(depth[::16, ::16], torch.from_numpy(gold["depth_1536"])).max()) <= 1e-4
    canon, fov = model_fp32.forward(x[None].to(DEV))
    assert canon.shape == (1, 1, 1536, 1536) and fov.shape == (1, 1, 1, 1)
    assert float(_pix_rel(canon[0, 0, ::16, ::16].cpu(), torch.from_numpy(gold["canon_1536"])).max()) <= 1e-4
    assert abs(float(fov) - float(gold["fov_deg_1536"][0])) / float(gold["fov_deg_1536"][0]) <= 1e-4


def test_fp32_1080p_u8_vs_golden(model_fp32, golden_dir):
    """uint8 frame -> fused transform + resize -> infer -> resize back (generate_depth_maps.py:113-121)."""
    gold = np.load(os.path.join(golden_dir, "reference_outputs.npz"))
    frame = O.synthetic_frame_u8(0)
    pred = model_fp32.infer(torch.from_numpy(frame))
    assert pred["depth"].shape == (1080, 1920)
    rel = _pix_rel(pred["depth"].cpu()[::16, ::16], torch.from_numpy(gold["depth_1080p"]))
    f_rel = abs(float(pred["focallength_px"]) - float(gold["f_px_1080p"])) / float(gold["f_px_1080p"])
    print(f"fp32 1080p: depth rel-err max {float(rel.max()):.3e}; f_px rel {f_rel:.3e}")
    assert float(rel.max()) <= 1e-4 and f_rel <= 1e-4
    # same frame through the reference-style transformed float tensor
    pred_f = model_fp32.infer(O.transform_u8(frame).to(DEV))
    assert torch.equal(pred_f["depth"], pred["depth"])
    # caller-supplied focal length (depth_pro.py:285-286)
    pred3 = model_fp32.infer(torch.from_numpy(frame), f_px=torch.tensor(1234.5))
    rel3 = _pix_rel(pred3["depth"].cpu()[::16, ::16], torch.from_numpy(gold["depth_1080p_fpx1234_5"]))
    assert float(rel3.max()) <= 1e-4
    assert float(pred3["focallength_px"]) == 1234.5


def test_bf16_vs_oracle(model_bf16, oracle_run):
    x, ref, taps = oracle_run
    pred = model_bf16.infer(x.to(DEV))
    depth = pred["depth"].cpu()
    ok = (ref["depth"] < 1e4 - 1) & (depth < 1e4 - 1)  # pixels not clamped on either side
    rel = _pix_rel(depth, ref["depth"])[ok]
    f_rel = abs(float(pred["focallength_px"]) - float(ref["focallength_px"])) / float(ref["focallength_px"])
    q = torch.quantile(rel[:: max(1, rel.numel() // 1_000_000)].float(), torch.tensor([0.5, 0.9, 0.99]))
    print(f"bf16: depth abs-rel median {float(q[0]):.3e} p90 {float(q[1]):.3e} p99 {float(q[2]):.3e} "
          f"max {float(rel.max()):.3e}; f_px rel {f_rel:.3e}; clamp mismatch {float((~ok).float().mean()):.2e}")
    for name in ("lat0_merged", "x0_merged", "global_tokens", "enc0", "enc4", "lowres", "decoder_out"):
        r = taps[name]
        got = model_bf16.tap(name).cpu().reshape(r.shape)
        print(f"  tap {name:14s} rms-err/rms {float((got - r).pow(2).mean().sqrt() / r.pow(2).mean().sqrt()):.3e}")
    assert float(q[0]) <= 5e-3
    assert f_rel <= 1e-2
    assert float((~ok).float().mean()) <= 1e-3
    # the tail is asserted too: p99 and max over ALL 2.36 M un-clamped pixels
    assert float(q[2]) <= 2e-2 and float(rel.max()) <= 5e-2
    # border ring: the composed head (head.1 o head.2, border-aware bias) and every padded conv differ from the
    # interior only on the outermost pixels -- same tolerance there as everywhere else
    ring = torch.zeros(1536, 1536, dtype=torch.bool)
    for i in (0, 1, 1534, 1535):
        ring[i, :] = True
        ring[:, i] = True
    rr = _pix_rel(depth, ref["depth"])[ring & (ref["depth"] < 1e4 - 1) & (depth < 1e4 - 1)].float()
    print(f"  border ring ({int(ring.sum())} px): median {float(rr.median()):.3e} p99 {float(torch.quantile(rr, 0.99)):.3e} "
          f"max {float(rr.max()):.3e}")
    assert float(rr.median()) <= 5e-3 and float(torch.quantile(rr, 0.99)) <= 2e-2 and float(rr.max()) <= 5e-2
    corners = _pix_rel(depth, ref["depth"])[[0, 0, 1535, 1535], [0, 1535, 0, 1535]]
    assert float(corners.max()) <= 5e-2


def test_bf16_1080p_u8_vs_golden(model_bf16, golden_dir):
    """The bf16 engine on the uint8 1080p video frame (fused transform + resize in, resize back out) against the
    REFERENCE's recorded fp32 output for that frame (tests/golden: depth_1080p, generate_depth_maps.py:113-121)."""
    gold = np.load(os.path.join(golden_dir, "reference_outputs.npz"))
    pred = model_bf16.infer(torch.from_numpy(O.synthetic_frame_u8(0)))
    assert pred["depth"].shape == (1080, 1920)
    g = torch.from_numpy(gold["depth_1080p"])
    d = pred["depth"].cpu()[::16, ::16]
    ok = (g < 1e4 - 1) & (d < 1e4 - 1)
    rel = _pix_rel(d, g)[ok].float()
    f_rel = abs(float(pred["focallength_px"]) - float(gold["f_px_1080p"])) / float(gold["f_px_1080p"])
    print(f"bf16 1080p: depth abs-rel median {float(rel.median()):.3e} p99 {float(torch.quantile(rel, 0.99)):.3e} "
          f"max {float(rel.max()):.3e}; f_px rel {f_rel:.3e}")
    assert float(rel.median()) <= 5e-3 and float(torch.quantile(rel, 0.99)) <= 2e-2 and float(rel.max()) <= 5e-2
    assert f_rel <= 1e-2 and float((~ok).float().mean()) <= 1e-3
    # last row / last column of the 1080p output (never on a ::16 grid of 1080 rows): finite and inside the clamp
    edge = torch.cat([pred["depth"][-1, :], pred["depth"][:, -1]])
    assert bool(torch.isfinite(edge).all()) and float(edge.min()) >= 1e-4 - 1e-9 and float(edge.max()) <= 1e4 * (1 + 1e-6)


def _oracle_vs_bf16(sd, seed_img):
    torch.set_num_threads(os.cpu_count())
    x = O.synthetic_image_1536(seed_img)
    ref = O.infer(sd, x)
    m = _model(sd, torch.bfloat16)
    pred = m.infer(x.to(DEV))
    canon, _ = m.forward(x[None].to(DEV))
    del m
    torch.cuda.empty_cache()
    return x, ref, pred, canon[0, 0].cpu()


def test_bf16_second_seed():
    """A second weight seed AND a second input: the bf16 claim must not rest on one draw."""
    sd = weights.stress_init(4321)
    _, ref, pred, _ = _oracle_vs_bf16(sd, 2)
    depth = pred["depth"].cpu()
    ok = (ref["depth"] < 1e4 - 1) & (depth < 1e4 - 1)
    rel = _pix_rel(depth, ref["depth"])[ok].float()
    q = torch.quantile(rel[:: max(1, rel.numel() // 1_000_000)], torch.tensor([0.5, 0.99]))
    f_rel = abs(float(pred["focallength_px"]) - float(ref["focallength_px"])) / float(ref["focallength_px"])
    # This draw has pixels whose inverse depth comes close to the final ReLU's kink (seed 1234 has none: min 0.74), where
    # a RELATIVE error is unbounded (measured max 0.22 on an inverse depth of 0.05).  The max is therefore asserted as
    # (a) relative error on pixels whose reference inverse depth is at least a quarter of the median, and (b) absolute
    # error of the inverse depth, in units of the median inverse depth, everywhere.
    inv_ref, inv = 1.0 / ref["depth"].double(), 1.0 / depth.double()
    scale = float(inv_ref[ok].median())
    solid = ok & (inv_ref >= 0.25 * scale)
    rel_solid = _pix_rel(depth, ref["depth"])[solid].float()
    abs_inv = ((inv - inv_ref).abs() / scale)[ok].float()
    print(f"bf16 seed 4321 / image 2: median {float(q[0]):.3e} p99 {float(q[1]):.3e} max {float(rel.max()):.3e} "
          f"(inverse depth >= median/4: {float(solid.float().mean()):.4f} of pixels, max {float(rel_solid.max()):.3e}); "
          f"|d inv| / median inv max {float(abs_inv.max()):.3e}; f_px {f_rel:.3e}")
    assert float(q[0]) <= 5e-3 and float(q[1]) <= 2e-2
    assert float(rel_solid.max()) <= 5e-2 and float(abs_inv.max()) <= 2e-2
    assert f_rel <= 1e-2 and float((~ok).float().mean()) <= 1e-3


def test_bf16_adversarial_zeros_init():
    """SURVEY.md §7 hard-part 1: an init whose final pre-ReLU map straddles zero (head.4 bias 0, weight x4), so a large
    share of the canonical inverse depth is EXACTLY 0 (depth clamped at 1e4) and the rest comes arbitrarily close to the
    ReLU kink.  Relative error is unbounded at the kink, so the grading is: (1) the zero masks agree except where the
    reference value itself is within bf16 noise of 0; (2) absolute error of the canonical inverse depth, relative to its
    scale, over ALL pixels; (3) abs-rel of depth on pixels comfortably above the kink."""
    sd = weights.stress_init(1234)
    sd["head.4.bias"] = torch.zeros(1)
    sd["head.4.weight"] = sd["head.4.weight"] * 4.0
    x = O.synthetic_image_1536(1)
    torch.set_num_threads(os.cpu_count())
    canon_ref, _ = O.forward(sd, x[None])
    canon_ref = canon_ref[0, 0]
    m = _model(sd, torch.bfloat16)
    canon, _ = m.forward(x[None].to(DEV))
    canon = canon[0, 0].cpu()
    del m
    zero_frac = float((canon_ref == 0).float().mean())
    scale = float(canon_ref[canon_ref > 0].median())
    abs_err = (canon - canon_ref).abs() / scale
    flips = (canon == 0) != (canon_ref == 0)
    flip_mag = torch.maximum(canon, canon_ref)[flips] / scale            # how far from the kink the disagreeing pixels are
    far = canon_ref > 0.5 * scale
    rel_far = ((canon - canon_ref).abs() / canon_ref)[far]
    print(f"adversarial: zeros {zero_frac:.2%}, scale {scale:.3f}; |err|/scale median {float(abs_err.median()):.3e} "
          f"max {float(abs_err.max()):.3e}; mask flips {float(flips.float().mean()):.2e} (largest {float(flip_mag.max()) if flips.any() else 0:.3e}); "
          f"rel on far pixels median {float(rel_far.median()):.3e} max {float(rel_far.max()):.3e}")
    assert 0.05 <= zero_frac <= 0.95, "the init is meant to put a large share of the map on the ReLU kink"
    # head.4's weight x4 with no bias makes the output a zero-mean sum: the same absolute bf16 noise of the 32 head
    # channels is 4x larger relative to the output than in recipe B (measured: median 5.3e-3, max 5.8e-2 of the scale).
    # These are robustness guards for the kink, not the BASELINE.md tolerance (that one is asserted on recipe B above).
    assert float(abs_err.max()) <= 1e-1 and float(abs_err.median()) <= 1e-2
    assert float(flips.float().mean()) <= 2e-2 and (not flips.any() or float(flip_mag.max()) <= 1e-1)
    assert float(rel_far.median()) <= 1e-2 and float(rel_far.max()) <= 2e-1


def test_bf16_layernorm_fold_with_outlier_channels(monkeypatch):
    """real DINOv2-L checkpoints carry a few residual-stream channels with activations two orders of
    magnitude above the rest.  The folded LayerNorm subtracts rstd*mean*colsum AFTER the GEMM on the bf16-rounded raw
    x, which is where such channels could cost accuracy: plant them (pos_embed + 40 in three channels of every encoder,
    ~100x the typical |x|), and require the folded path to stay as close to the fp32 oracle as the stand-alone one."""
    sd = weights.stress_init(1234)
    for p in weights.VIT_PREFIXES:
        pe = sd[p + "pos_embed"].clone()
        pe[..., [5, 300, 911]] += torch.tensor([40.0, -55.0, 70.0])
        sd[p + "pos_embed"] = pe
    x = O.synthetic_image_1536(1)
    torch.set_num_threads(os.cpu_count())
    ref = O.infer(sd, x)["depth"]
    folded_m = _model(sd, torch.bfloat16)
    folded = folded_m.infer(x.to(DEV))["depth"].cpu()
    del folded_m
    monkeypatch.setenv("DEPTHPRO_LN_FUSE", "0")
    plain_m = _model(sd, torch.bfloat16)
    plain = plain_m.infer(x.to(DEV))["depth"].cpu()     # the engine (and its LN mode) is created by the first call
    monkeypatch.delenv("DEPTHPRO_LN_FUSE")
    del plain_m
    assert not torch.equal(plain, folded), "DEPTHPRO_LN_FUSE=0 did not select the stand-alone LayerNorm path"
    ok = (ref < 1e4 - 1) & (folded < 1e4 - 1) & (plain < 1e4 - 1)
    e_f, e_p = _pix_rel(folded, ref)[ok].float(), _pix_rel(plain, ref)[ok].float()
    print(f"outlier channels: folded median {float(e_f.median()):.3e} p99 {float(torch.quantile(e_f[::4], 0.99)):.3e}; "
          f"stand-alone median {float(e_p.median()):.3e} p99 {float(torch.quantile(e_p[::4], 0.99)):.3e}")
    assert float(e_f.median()) <= 5e-3
    assert float(e_f.median()) <= 1.5 * float(e_p.median()) + 5e-4


def test_bf16_head0_composed_with_decoder_out_conv(model_bf16, oracle_run, monkeypatch):
    """Default bf16 engine: `decoder.fusions.0.out_conv` (1x1, decoder.py:178) is composed into `head.0` (conv3x3,
    depth_pro.py:183-185) at finalize, with a border-aware bias (the 1x1's bias does not exist in head.0's zero padding).
    DEPTHPRO_HEAD0_FUSE=0 keeps the two launches.  One launch less per frame, the same depth far inside the bf16
    tolerance -- on the outermost pixel ring as well, which is where a wrong border term would show -- and the
    `decoder_out` tap (now evaluated on demand) still matches the oracle."""
    x, ref, taps = oracle_run
    fused = model_bf16.infer(x.to(DEV))
    n0 = model_bf16.launch_count()
    model_bf16.infer(x.to(DEV))
    per_frame_fused = model_bf16.launch_count() - n0
    got = model_bf16.tap("decoder_out").cpu().reshape(taps["decoder_out"].shape)
    r = taps["decoder_out"]
    assert float((got - r).abs().max() / r.abs().max()) < 3e-2
    monkeypatch.setenv("DEPTHPRO_HEAD0_FUSE", "0")
    plain_model = depth_pro.DepthPro(device=DEV, precision=torch.bfloat16).init_weights("stress", 1234).eval()
    plain = plain_model.infer(x.to(DEV))
    monkeypatch.delenv("DEPTHPRO_HEAD0_FUSE")
    n0 = plain_model.launch_count()
    plain_model.infer(x.to(DEV))
    assert plain_model.launch_count() - n0 == per_frame_fused + 1
    d_a, d_b, d_r = fused["depth"].cpu(), plain["depth"].cpu(), ref["depth"]
    assert not torch.equal(d_a, d_b)
    ok = (d_r < 1e4 - 1) & (d_a < 1e4 - 1) & (d_b < 1e4 - 1)
    rel = _pix_rel(d_a, d_b)
    ring = torch.zeros_like(ok)
    ring[:2], ring[-2:], ring[:, :2], ring[:, -2:] = True, True, True, True
    e_in, e_ring = rel[ok & ~ring].float(), rel[ok & ring].float()
    print(f"head.0 composed vs separate: interior median {float(e_in.median()):.3e}, border ring median "
          f"{float(e_ring.median()):.3e} max {float(e_ring.max()):.3e}")
    assert float(e_in.median()) <= 1.5e-3
    assert float(e_ring.median()) <= 2.0 * float(e_in.median()) + 5e-4 and float(e_ring.max()) <= 5e-2
    e_a, e_b = _pix_rel(d_a, d_r)[ok].float().median(), _pix_rel(d_b, d_r)[ok].float().median()
    assert float(e_a) <= 5e-3 and float(e_a) <= 1.2 * float(e_b) + 2e-4
    del plain_model


def test_bf16_serpentine_tile_order_is_bit_identical(model_bf16, oracle_run, monkeypatch):
    """Default: the ViT kernels alternate their tile direction (`GemmOp::reverse`, attention units likewise) so that a
    consumer starts on what the L2 still holds of its producer's output.  Tiles are independent and the LayerNorm
    partial sums are added in a fixed slot order, so DEPTHPRO_SERPENTINE=0 (read at engine creation) must give the same
    bits."""
    x, _, _ = oracle_run
    a = model_bf16.infer(x.to(DEV))
    monkeypatch.setenv("DEPTHPRO_SERPENTINE", "0")
    plain_model = depth_pro.DepthPro(device=DEV, precision=torch.bfloat16).init_weights("stress", 1234).eval()
    b = plain_model.infer(x.to(DEV))
    monkeypatch.delenv("DEPTHPRO_SERPENTINE")
    assert torch.equal(a["depth"], b["depth"]) and float(a["focallength_px"]) == float(b["focallength_px"])
    del plain_model


def test_bf16_pair_residual_stream_matches_fp32_stream(model_bf16, oracle_run, monkeypatch):
    """Default: the ViT residual stream is a (hi, lo) pair of 16-bit arrays whose hi half is the next GEMM's operand
    (csrc/common.cuh GemmOp::ln_xlo).  DEPTHPRO_RES_PAIR=0 (read at engine creation) keeps the fp32 stream + separate
    16-bit copy.  The pair resolves 2^-16 per value against the 2^-9 of every GEMM operand: the two engines must agree
    far inside the bf16 tolerance and sit equally close to the fp32 reference."""
    x, ref, _ = oracle_run
    pair = model_bf16.infer(x.to(DEV))
    monkeypatch.setenv("DEPTHPRO_RES_PAIR", "0")
    f32_model = depth_pro.DepthPro(device=DEV, precision=torch.bfloat16).init_weights("stress", 1234).eval()
    f32 = f32_model.infer(x.to(DEV))
    monkeypatch.delenv("DEPTHPRO_RES_PAIR")
    d_a, d_b, d_r = pair["depth"].cpu(), f32["depth"].cpu(), ref["depth"]
    assert not torch.equal(d_a, d_b), "DEPTHPRO_RES_PAIR=0 did not select the fp32 residual stream"
    ok = (d_r < 1e4 - 1) & (d_a < 1e4 - 1) & (d_b < 1e4 - 1)
    between = _pix_rel(d_a, d_b)[ok].float()
    e_a, e_b = _pix_rel(d_a, d_r)[ok].float().median(), _pix_rel(d_b, d_r)[ok].float().median()
    print(f"pair vs fp32 residual stream: median rel diff {float(between.median()):.3e}; vs fp32 reference: pair "
          f"{float(e_a):.3e}, fp32 stream {float(e_b):.3e}")
    assert float(between.median()) <= 1.5e-3
    assert float(e_a) <= 5e-3 and float(e_a) <= 1.2 * float(e_b) + 2e-4
    assert abs(float(pair["focallength_px"]) - float(f32["focallength_px"])) <= 1e-3 * float(f32["focallength_px"])
    del f32_model


def test_bf16_layernorm_fold_matches_standalone_layernorm(model_bf16, oracle_run, monkeypatch):
    """The LayerNorm-folded ViT (default) and the stand-alone LayerNorm launches (DEPTHPRO_LN_FUSE=0, read at
    engine creation) are two roundings of the same arithmetic: they must agree far inside the bf16 tolerance,
    and be equally close to the fp32 reference."""
    x, ref, _ = oracle_run
    folded = model_bf16.infer(x.to(DEV))
    monkeypatch.setenv("DEPTHPRO_LN_FUSE", "0")
    plain_model = depth_pro.DepthPro(device=DEV, precision=torch.bfloat16).init_weights("stress", 1234).eval()
    plain = plain_model.infer(x.to(DEV))
    monkeypatch.delenv("DEPTHPRO_LN_FUSE")
    n0 = plain_model.launch_count()
    plain_model.infer(x.to(DEV))
    per_frame_plain = plain_model.launch_count() - n0
    n0 = model_bf16.launch_count()
    model_bf16.infer(x.to(DEV))
    per_frame_folded = model_bf16.launch_count() - n0
    assert per_frame_plain - per_frame_folded == 48 - 1     # 48 LayerNorm launches gone, one statistics pass added
    d_f, d_p, d_r = folded["depth"].cpu(), plain["depth"].cpu(), ref["depth"]
    ok = (d_r < 1e4 - 1) & (d_f < 1e4 - 1) & (d_p < 1e4 - 1)
    between = _pix_rel(d_f, d_p)[ok].float()
    e_f, e_p = _pix_rel(d_f, d_r)[ok].float().median(), _pix_rel(d_p, d_r)[ok].float().median()
    print(f"LN fold vs stand-alone: median rel diff {float(between.median()):.3e}; vs fp32 reference: folded "
          f"{float(e_f):.3e}, stand-alone {float(e_p):.3e}")
    assert float(between.median()) <= 3e-3
    assert float(e_f) <= 5e-3 and float(e_f) <= 1.5 * float(e_p) + 5e-4
    del plain_model


def test_fp16_mode_vs_oracle(state_dict, oracle_run, model_bf16, golden_dir):
    """VERDICT r1 missing #4: `precision=torch.half` (the reference's model.half(), depth_pro.py:122-123) runs a TRUE fp16
    engine (libdepthpro_b200_fp16.so: the same sources with IEEE-half storage, fp32 accumulation / residual stream /
    statistics).  Against the fp32 oracle it must meet the bf16 tolerance with room to spare -- the reference's own
    all-fp16 arithmetic measures median 3.3e-4 on this frame (profiles/r2_reduced_precision_probe.json) -- and it must
    really be a different arithmetic from the bf16 engine."""
    x, ref, taps = oracle_run
    m = _model(state_dict, torch.float16)
    assert m._lib_flavour == "fp16"
    pred = m.infer(x.to(DEV))
    depth = pred["depth"].cpu()
    ok = (ref["depth"] < 1e4 - 1) & (depth < 1e4 - 1)
    rel = _pix_rel(depth, ref["depth"])[ok].float()
    q = torch.quantile(rel[:: max(1, rel.numel() // 1_000_000)], torch.tensor([0.5, 0.99]))
    f_rel = abs(float(pred["focallength_px"]) - float(ref["focallength_px"])) / float(ref["focallength_px"])
    print(f"fp16: depth abs-rel median {float(q[0]):.3e} p99 {float(q[1]):.3e} max {float(rel.max()):.3e}; f_px rel {f_rel:.3e}")
    for name in ("lat0_merged", "x0_merged", "enc0", "enc4", "decoder_out"):
        r = taps[name]
        got = m.tap(name).cpu().reshape(r.shape)
        print(f"  tap {name:14s} rms-err/rms {float((got - r).pow(2).mean().sqrt() / r.pow(2).mean().sqrt()):.3e}")
    assert float(q[0]) <= 1e-3 and float(q[1]) <= 5e-3 and float(rel.max()) <= 2e-2
    assert f_rel <= 1e-3 and float((~ok).float().mean()) <= 1e-3
    bf = model_bf16.infer(x.to(DEV))["depth"].cpu()
    assert not torch.equal(bf, depth)
    assert float(_pix_rel(bf, ref["depth"])[ok].float().median()) > float(q[0])      # fp16 is the closer one here
    # uint8 video frame through the fp16 engine vs the reference's recorded output
    gold = np.load(os.path.join(golden_dir, "reference_outputs.npz"))
    p2 = m.infer(torch.from_numpy(O.synthetic_frame_u8(0)))
    r2 = _pix_rel(p2["depth"].cpu()[::16, ::16], torch.from_numpy(gold["depth_1080p"])).float()
    assert float(r2.median()) <= 1e-3 and float(r2.max()) <= 2e-2
    del m
    torch.cuda.empty_cache()


def test_interpolation_mode_bicubic(model_fp32, state_dict):
    """VERDICT r1 missing #2: DepthPro.infer(interpolation_mode="bicubic") (depth_pro.py:247, 273-279, 288-291) against the
    oracle in fp32, down- and up-sampling in one call (540x960 -> 1536^2 -> 540x960), tolerance of the fp32 mode."""
    x = O.transform_u8(O.synthetic_frame_u8(2, 540, 960))
    torch.set_num_threads(os.cpu_count())
    ref = O.infer(state_dict, x, interpolation_mode="bicubic")
    pred = model_fp32.infer(x.to(DEV), interpolation_mode="bicubic")
    rel = _pix_rel(pred["depth"].cpu(), ref["depth"])
    f_rel = abs(float(pred["focallength_px"]) - float(ref["focallength_px"])) / float(ref["focallength_px"])
    print(f"fp32 bicubic 540x960: depth rel-err max {float(rel.max()):.3e}; f_px rel {f_rel:.3e}")
    assert float(rel.max()) <= 1e-4 and f_rel <= 1e-4
    bil = model_fp32.infer(x.to(DEV))["depth"].cpu()
    assert float(_pix_rel(bil, ref["depth"]).max()) > 1e-3          # the mode really changes the result


@pytest.mark.parametrize("fov", [None, "head"])
def test_non_default_fov_configs(fov):
    """VERDICT r1 missing #3: `use_fov_head=False` and `fov_encoder_preset=None` (depth_pro.py:100-108, 236-241;
    fov.py:29-56) through create_model_and_transforms, against the oracle (whose conv-only head is pinned to the
    reference's own FOVNetwork in tests/test_oracle.py)."""
    import dataclasses

    cfg = dataclasses.replace(depth_pro.depth_pro.DEFAULT_MONODEPTH_CONFIG_DICT, checkpoint_uri=None,
                              use_fov_head=fov is not None, fov_encoder_preset=None)
    model, _ = depth_pro.create_model_and_transforms(cfg, device=DEV, precision=torch.bfloat16)
    sd = weights.stress_init(SEED, fov=fov)
    assert set(model.state_dict()) == set(sd)
    model.load_state_dict(sd, strict=True)
    x = O.synthetic_image_1536(1)
    torch.set_num_threads(os.cpu_count())
    canon_ref, fov_ref = O.forward(sd, x[None])
    canon, fov_deg = model.forward(x[None].to(DEV))
    rel = _pix_rel(canon[0, 0].cpu(), canon_ref[0, 0]).float()
    print(f"fov={fov}: canonical inverse depth abs-rel median {float(rel.median()):.3e} max {float(rel.max()):.3e}")
    assert float(rel.median()) <= 5e-3 and float(rel.max()) <= 5e-2
    if fov is None:
        assert fov_deg is None and fov_ref is None                   # depth_pro.py:236-241
        with pytest.raises(AttributeError):                          # depth_pro.py:282-283 on fov_deg=None
            model.infer(x.to(DEV))
    else:
        assert fov_deg.shape == (1, 1, 1, 1)
        assert abs(float(fov_deg) - float(fov_ref)) / abs(float(fov_ref)) <= 1e-2
        est = model.infer(x.to(DEV))
        ref = O.infer(sd, x)
        assert abs(float(est["focallength_px"]) - float(ref["focallength_px"])) / float(ref["focallength_px"]) <= 1e-2
    pred = model.infer(x.to(DEV), f_px=1500.0)                       # caller-supplied focal length works in every config
    ref = O.infer(sd, x, f_px=1500.0)
    r2 = _pix_rel(pred["depth"].cpu(), ref["depth"]).float()
    assert float(r2.median()) <= 5e-3 and float(pred["focallength_px"]) == 1500.0
    del model
    torch.cuda.empty_cache()


def test_weight_edits_reach_the_engine(state_dict):
    """module-level conversions re-upload automatically, in-place edits need refresh_weights()."""
    m = _model(state_dict, torch.bfloat16)
    x = O.synthetic_image_1536(1)[:, :256, :256].contiguous().to(DEV)
    a = m.infer(x)["depth"].clone()
    with torch.no_grad():
        m.head.get_parameter("4.bias").add_(0.5)
    assert torch.equal(m.infer(x)["depth"], a)                       # documented: the packed copy is still in use
    m.refresh_weights()
    b = m.infer(x)["depth"].clone()
    assert not torch.equal(b, a)
    m.float()                                                        # nn.Module._apply marks the engine copy stale
    assert m._dirty
    assert torch.equal(m.infer(x)["depth"], b)
    del m


def test_two_engines_on_two_threads(state_dict, model_bf16):
    """the process-wide pieces of the library (tensor-map cache, launch counter, profiler,
    A/B switches) are guarded, so two engines driven from two host threads on their own streams may run concurrently
    (SURVEY.md §8e "one host thread per GPU").  Each thread's results must equal the single-threaded ones bit for bit."""
    import threading

    frames = [torch.from_numpy(O.synthetic_frame_u8(i, 270, 480)).to(DEV) for i in range(6)]
    want = [model_bf16.infer(f)["depth"].clone() for f in frames]
    second = _model(state_dict, torch.bfloat16)
    models, got, errors = [model_bf16, second], [[None] * 6, [None] * 6], []
    start = threading.Barrier(2)

    def work(t):
        try:
            stream = torch.cuda.Stream(DEV)
            with torch.cuda.stream(stream):
                start.wait()
                for rep in range(2):                 # eager call, graph capture, replays -- interleaved across threads
                    for i in range(6):
                        j = i if t == 0 else 5 - i
                        got[t][j] = models[t].infer(frames[j])["depth"].clone()
                stream.synchronize()
        except Exception as e:  # noqa: BLE001
            errors.append(repr(e))

    threads = [threading.Thread(target=work, args=(t,)) for t in range(2)]
    [t.start() for t in threads]
    [t.join() for t in threads]
    assert not errors, errors
    for t in range(2):
        for i in range(6):
            assert torch.equal(got[t][i], want[i]), (t, i)
    del second
    torch.cuda.empty_cache()


def test_infer_host_entry_point(model_bf16):
    """`dp_infer_host` (include/depthpro_b200.h: the call generate_depth_maps.py:113-121 amounts to -- host frame in,
    host depth out, one synchronous C call) equals `model.infer` on the same frame bit for bit, with and without a
    caller-supplied focal length."""
    import ctypes

    from depth_pro import _capi

    frame = np.ascontiguousarray(O.synthetic_frame_u8(3, 270, 480))
    lib = model_bf16._ensure_engine(1)
    depth = np.empty((270, 480), np.float32)
    f_out = np.empty((1,), np.float32)
    _capi.check(lib.dp_infer_host(model_bf16._engine, frame.ctypes.data, 1, 270, 480, _capi.SRC_U8_HWC, None,
                                  depth.ctypes.data, f_out.ctypes.data), lib)
    want = model_bf16.infer(torch.from_numpy(frame))
    assert np.array_equal(depth, want["depth"].cpu().numpy()) and float(f_out[0]) == float(want["focallength_px"])
    f_in = np.array([777.0], np.float32)
    _capi.check(lib.dp_infer_host(model_bf16._engine, frame.ctypes.data, 1, 270, 480, _capi.SRC_U8_HWC, f_in.ctypes.data,
                                  depth.ctypes.data, f_out.ctypes.data), lib)
    want = model_bf16.infer(torch.from_numpy(frame), f_px=777.0)
    assert np.array_equal(depth, want["depth"].cpu().numpy()) and float(f_out[0]) == 777.0


def test_batch_is_bit_identical(model_bf16):
    """Frames are independent units: a 2-frame batch must equal two single-frame calls bit for bit."""
    frames = np.stack([O.synthetic_frame_u8(i, 540, 960) for i in range(2)])
    both = model_bf16.infer(torch.from_numpy(frames))
    for i in range(2):
        one = model_bf16.infer(torch.from_numpy(frames[i]))
        assert torch.equal(both["depth"][i], one["depth"])
        assert torch.equal(both["focallength_px"][i], one["focallength_px"])


def test_errors():
    with pytest.raises(RuntimeError):
        depth_pro.create_model_and_transforms(device=torch.device("cpu"))
    with pytest.raises(KeyError):
        depth_pro.depth_pro.create_backbone_model("nope")


def test_infer_edge_shapes_and_argument_forms(model_bf16):
    """Ragged / extreme inputs of DepthPro.infer (depth_pro.py:243-298): tiny and odd image sizes, 3-D vs 4-D input,
    non-contiguous tensors, every accepted form of f_px, and the error paths."""
    rng = np.random.default_rng(9)
    for H, W in ((1, 1), (1, 7), (5, 3), (17, 23), (1536, 1536), (2001, 333)):
        img = rng.integers(0, 256, size=(H, W, 3), dtype=np.uint8)
        pred = model_bf16.infer(img)                          # ndarray straight from load_rgb
        d = pred["depth"]
        assert d.shape == ((H, W) if H > 1 and W > 1 else torch.Size([s for s in (H, W) if s > 1]))   # reference squeezes
        assert d.dtype == torch.float32 and d.is_cuda and bool(torch.isfinite(d).all())
        assert float(d.min()) >= 1e-4 - 1e-9 and float(d.max()) <= 1e4 * (1 + 1e-6)                  # clamp(1e-4, 1e4) of the inverse
        assert pred["focallength_px"].dim() == 0 and float(pred["focallength_px"]) > 0
    # float CHW in [-1, 1] (the reference's transformed tensor), 3-D and 4-D, contiguous or not: same result
    x = torch.rand(3, 40, 56) * 2 - 1
    a = model_bf16.infer(x)["depth"]
    b = model_bf16.infer(x[None])["depth"]
    c = model_bf16.infer(x.permute(1, 2, 0).contiguous().permute(2, 0, 1))["depth"]     # non-contiguous view
    assert a.shape == (40, 56) and torch.equal(a, b) and torch.equal(a, c)
    assert torch.equal(model_bf16.infer(x.double())["depth"], a)                        # any float dtype is accepted
    # f_px: python float, 0-d tensor, 1-element tensor, per-image tensor; the value is passed through (:285-286)
    for f in (500.0, torch.tensor(500.0), torch.tensor([500.0])):
        p = model_bf16.infer(x, f_px=f)
        assert float(p["focallength_px"]) == 500.0
    two = torch.stack([x, x.flip(-1)])
    p2 = model_bf16.infer(two, f_px=torch.tensor([400.0, 800.0]))
    assert p2["depth"].shape == (2, 40, 56)
    one = model_bf16.infer(x, f_px=400.0)["depth"]
    assert torch.equal(p2["depth"][0], one)
    # depth scales with W / f_px: doubling f_px doubles metric depth wherever neither hits the clamp
    far = model_bf16.infer(x, f_px=800.0)["depth"]
    ok = (one < 4e3) & (one > 1e-3)
    assert float(((far / one)[ok] - 2.0).abs().max()) < 1e-5
    # interpolation_mode: F.interpolate(align_corners=False) accepts bilinear / bicubic only (reference: ValueError)
    for bad in ("nearest", "area", "nearest-exact", "trilinear"):
        with pytest.raises(ValueError):
            model_bf16.infer(x, interpolation_mode=bad)
    assert not torch.equal(model_bf16.infer(x, interpolation_mode="bicubic")["depth"], a)
    with pytest.raises(AssertionError):
        model_bf16.infer(torch.rand(4, 40, 56))               # not 3 channels
    with pytest.raises(AssertionError):
        model_bf16.forward(torch.rand(1, 3, 384, 384))        # forward is 1536^2 only (:231)
    with pytest.raises(AssertionError):
        model_bf16.infer(x, f_px=torch.tensor([1.0, 2.0, 3.0]))
