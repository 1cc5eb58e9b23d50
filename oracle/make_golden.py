"""TEST INFRASTRUCTURE ONLY — generate tests/golden/* from the UNMODIFIED reference.

Run in the build container (needs /root/reference):  python oracle/make_golden.py
It (1) dumps the reference's state_dict manifest, (2) records the index maps produced by
the reference's own `DepthProEncoder.split/merge` on index-coded tensors, (3) runs the
reference `model.infer` (CPU fp32, "recipe B" seed 1234) on the Config-1 1536^2 input and
on a 1080p synthetic frame, checks the oracle restatement against it tensor-by-tensor,
and stores strided samples + statistics of the outputs and stage taps, and (4) executes the
reference's `depth_to_3d` source on a small depth map with NaN / <=0 entries.
The fixtures are small (< 1 MB total) and committed; the GPU box never needs the reference.
"""

from __future__ import annotations

import ast
import hashlib
import json
import os
import sys
import time

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, HERE)
sys.path.insert(0, os.path.join(ROOT, "ml-depth-pro-video_b200"))

import depthpro_oracle as O  # noqa: E402
import reference_loader as RL  # noqa: E402
from depth_pro import weights as W  # noqa: E402

GOLD = os.path.join(ROOT, "tests", "golden")
STRIDE = 16


def stats(t: torch.Tensor) -> dict:
    t = t.detach().float()
    return {"shape": list(t.shape), "mean": float(t.mean()), "std": float(t.std()),
            "absmax": float(t.abs().max())}


def relerr(a: torch.Tensor, b: torch.Tensor) -> float:
    return float((a - b).abs().max() / b.abs().max().clamp_min(1e-30))


def index_fixtures(ref_encoder_cls):
    """Bit-exact pins for split (encoder.py:170-188) and merge (encoder.py:190-217)."""
    out = {}
    # split: index-coded single-channel images; record each patch's corner + a digest.
    for name, size, ov in (("split_1536", 1536, 0.25), ("split_768", 768, 0.5)):
        x = torch.arange(size * size, dtype=torch.int32).reshape(1, 1, size, size)
        p = ref_encoder_cls.split(None, x, overlap_ratio=ov)
        out[name + "_corner"] = p[:, 0, 0, 0].numpy().copy()           # flat index of (j0, i0)
        out[name + "_last"] = p[:, 0, -1, -1].numpy().copy()
        out[name + "_sha256"] = np.frombuffer(
            hashlib.sha256(p.numpy().tobytes()).digest(), dtype=np.uint8).copy()
        # batch 2: patch-major / batch-minor ordering
        x2 = torch.stack([x[0], x[0] + size * size])
        p2 = ref_encoder_cls.split(None, x2, overlap_ratio=ov)
        out[name + "_b2_corner"] = p2[:, 0, 0, 0].numpy().copy()
    # merge: value = patch * 576 + token index; output map is small, keep it whole.
    for name, steps, pad in (("merge_5x5_pad3", 5, 3), ("merge_3x3_pad6", 3, 6)):
        n = steps * steps
        x = torch.arange(n * 576, dtype=torch.int32).reshape(n, 1, 24, 24)
        out[name] = ref_encoder_cls.merge(None, x, batch_size=1, padding=pad)[0, 0].numpy().copy()
        x2 = torch.arange(2 * n * 576, dtype=torch.int32).reshape(2 * n, 1, 24, 24)
        out[name + "_b2"] = ref_encoder_cls.merge(None, x2, batch_size=2, padding=pad)[:, 0].numpy().copy()
    np.savez_compressed(os.path.join(GOLD, "split_merge_index.npz"), **out)
    print("wrote split_merge_index.npz")


def depth_to_3d_fixture():
    """Execute the reference's own depth_to_3d (img_to_normalized_pointcloud.py:819-856)."""
    src_path = "/root/reference/img_to_normalized_pointcloud.py"
    tree = ast.parse(open(src_path).read())
    fn = next(n for n in tree.body if isinstance(n, ast.FunctionDef) and n.name == "depth_to_3d")
    ns = {"np": np}
    exec(compile(ast.Module(body=[fn], type_ignores=[]), src_path, "exec"), ns)
    rng = np.random.default_rng(5)
    h, w = 37, 53
    depth = rng.uniform(0.3, 40.0, size=(h, w)).astype(np.float32)
    depth[rng.uniform(size=(h, w)) < 0.05] = np.nan
    depth[rng.uniform(size=(h, w)) < 0.05] = 0.0
    depth[rng.uniform(size=(h, w)) < 0.05] = -1.0
    f = 41.7
    pts, valid = ns["depth_to_3d"](depth, f, w, h)
    np.savez_compressed(os.path.join(GOLD, "depth_to_3d.npz"), depth=depth, f=np.float64(f),
                        points=pts, valid=valid)
    o_pts, o_valid = O.depth_to_3d(depth, f, w, h)
    assert np.array_equal(o_valid, valid) and np.array_equal(o_pts, pts)
    print("wrote depth_to_3d.npz", pts.shape, pts.dtype)


def ground_fixture():
    """Execute the reference's own normalize_point_cloud_to_ground / grid_based_ground_adjustment
    (img_to_normalized_pointcloud.py:858-1118) on a seeded synthetic room and pin the oracle to them."""
    import contextlib
    import io

    src_path = "/root/reference/img_to_normalized_pointcloud.py"
    tree = ast.parse(open(src_path).read())
    want = ("point_plane_distances", "normalize_point_cloud_to_ground", "grid_based_ground_adjustment")
    fns = [n for n in tree.body if isinstance(n, ast.FunctionDef) and n.name in want]
    assert len(fns) == 3
    ns = {"np": np}
    exec(compile(ast.Module(body=fns, type_ignores=[]), src_path, "exec"), ns)
    out = {}
    for tag, tilt, seed in (("tilt12", 12.0, 3), ("tilt3", 3.0, 4)):   # tilt3: |normal.y| > 0.99, no rotation branch
        pts32, normal, d = O.synthetic_room_points(20000, seed, tilt)
        pts = pts32.astype(np.float64)
        with contextlib.redirect_stdout(io.StringIO()):
            norm = ns["normalize_point_cloud_to_ground"](pts, {"normal": normal, "d": d})
            # the second stage sees float32 points (what the GPU path hands over), widened again
            norm32 = norm.astype(np.float32)
            grid = ns["grid_based_ground_adjustment"](norm32.astype(np.float64), grid_size=20, percentile=5)
        o_norm = O.normalize_point_cloud_to_ground(pts, normal, d)
        o_grid = O.grid_based_ground_adjustment(norm32.astype(np.float64), 20, 5)
        assert np.max(np.abs(o_norm - norm)) <= 1e-12, np.max(np.abs(o_norm - norm))
        assert np.array_equal(o_grid, grid), np.max(np.abs(o_grid - grid))
        print(f"ground fixture {tag}: normalize max diff {np.max(np.abs(o_norm - norm)):.2e}; grid adjusted "
              f"{int((grid[:, 1] != norm32[:, 1]).sum())} points")
        out[tag + "_points"], out[tag + "_normal"], out[tag + "_d"] = pts32, normal, np.float64(d)
        out[tag + "_normalized"], out[tag + "_grid"] = norm.astype(np.float32), grid.astype(np.float32)
    np.savez_compressed(os.path.join(GOLD, "ground_normalize.npz"), **out)
    print("wrote ground_normalize.npz")


def main():
    os.makedirs(GOLD, exist_ok=True)
    torch.set_num_threads(os.cpu_count())
    ref = RL.load()
    enc_cls = sys.modules["ref_depth_pro.network.encoder"].DepthProEncoder
    index_fixtures(enc_cls)
    depth_to_3d_fixture()
    ground_fixture()

    t0 = time.time()
    sd = W.stress_init(1234)
    print(f"stress_init: {time.time() - t0:.1f}s")
    model, transform = RL.build_reference_model()
    man = {k: list(v.shape) for k, v in model.state_dict().items()}
    json.dump(man, open(os.path.join(GOLD, "state_dict_manifest.json"), "w"), indent=0)
    assert list(man) == list(W.manifest()) and all(tuple(man[k]) == tuple(s) for k, s in W.manifest().items())
    model.load_state_dict(sd, strict=True)

    taps_ref = {}
    model.encoder.register_forward_hook(lambda m, i, o: taps_ref.update(
        {f"enc{k}": v for k, v in enumerate(o)},
        hook0=m.backbone_highres_hook0, hook1=m.backbone_highres_hook1))
    model.decoder.register_forward_hook(lambda m, i, o: taps_ref.update(decoder_out=o[0], lowres=o[1]))
    model.head.register_forward_hook(lambda m, i, o: taps_ref.update(canonical_inverse_depth=o))

    gold = {}
    meta = {"seed": 1234, "stride": STRIDE, "taps": {}}

    # ---- Config 1: 1536^2 float input ------------------------------------------------
    x = O.synthetic_image_1536(1)
    t0 = time.time()
    with torch.no_grad():
        pred = model.infer(x)
        canon_ref, fov_ref = model.forward(x.unsqueeze(0))
    meta["ref_infer_plus_forward_s"] = time.time() - t0
    print(f"reference infer+forward: {time.time() - t0:.1f}s")
    t0 = time.time()
    taps_o = {}
    pred_o = O.infer(sd, x, taps=taps_o)
    meta["oracle_infer_s"] = time.time() - t0
    print(f"oracle infer: {time.time() - t0:.1f}s")

    report = {}
    for k in ("enc0", "enc1", "enc2", "enc3", "enc4", "decoder_out", "lowres", "canonical_inverse_depth"):
        report[k] = relerr(taps_o[k], taps_ref[k])
    report["fov_deg"] = relerr(taps_o["fov_deg"], fov_ref)
    report["depth"] = relerr(pred_o["depth"], pred["depth"])
    report["f_px"] = relerr(pred_o["focallength_px"], pred["focallength_px"])
    print("oracle vs reference (max abs err / absmax):", json.dumps(report, indent=1))
    meta["oracle_vs_reference"] = report
    assert max(report.values()) < 2e-5, report

    gold["depth_1536"] = pred["depth"][::STRIDE, ::STRIDE].numpy().copy()
    gold["f_px_1536"] = pred["focallength_px"].numpy().copy()
    gold["fov_deg_1536"] = fov_ref.reshape(-1).numpy().copy()
    gold["canon_1536"] = canon_ref[0, 0, ::STRIDE, ::STRIDE].numpy().copy()
    for k in ("enc0", "enc1", "enc2", "enc3", "enc4", "decoder_out", "lowres", "hook0", "hook1"):
        t = taps_ref[k]
        meta["taps"][k] = stats(t)
        if t.dim() == 4:  # NCHW: keep 8 channels on a coarse grid
            s = max(1, t.shape[-1] // 24)
            gold["tap_" + k] = t[0, :: t.shape[1] // 8, ::s, ::s].numpy().copy()
        else:  # tokens (35,577,1024)
            gold["tap_" + k] = t[::6, ::48, ::64].numpy().copy()
    inv = taps_ref["canonical_inverse_depth"]
    meta["canon_stats"] = dict(stats(inv), median=float(inv.median()), min=float(inv.min()),
                               zero_frac=float((inv == 0).float().mean()))
    print("canonical inverse depth:", meta["canon_stats"], "fov", float(fov_ref))

    # ---- Config 3 style: 1080p uint8 frame through transform + resize ------------------
    frame = O.synthetic_frame_u8(0)
    xt = transform(frame)
    t0 = time.time()
    with torch.no_grad():
        pred2 = model.infer(xt)
    print(f"reference infer 1080p: {time.time() - t0:.1f}s")
    pred2_o = O.infer(sd, O.transform_u8(frame))
    r2 = {"depth": relerr(pred2_o["depth"], pred2["depth"]),
          "f_px": relerr(pred2_o["focallength_px"], pred2["focallength_px"])}
    print("oracle vs reference 1080p:", r2)
    meta["oracle_vs_reference_1080p"] = r2
    assert max(r2.values()) < 2e-5, r2
    gold["depth_1080p"] = pred2["depth"][::STRIDE, ::STRIDE].numpy().copy()
    gold["f_px_1080p"] = pred2["focallength_px"].numpy().copy()
    # user-supplied focal length path (depth_pro.py:285-286)
    with torch.no_grad():
        pred3 = model.infer(xt, f_px=torch.tensor(1234.5))
    gold["depth_1080p_fpx1234_5"] = pred3["depth"][::STRIDE, ::STRIDE].numpy().copy()

    np.savez_compressed(os.path.join(GOLD, "reference_outputs.npz"), **gold)
    json.dump(meta, open(os.path.join(GOLD, "reference_outputs.json"), "w"), indent=1)
    print("wrote reference_outputs.npz / .json")


if __name__ == "__main__":
    main()
