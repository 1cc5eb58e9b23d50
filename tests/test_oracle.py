"""CPU tests of the oracle (the checker): golden vectors recorded from the unmodified reference,
plus a live comparison with the reference itself when /root/reference is present."""

import hashlib
import os
import sys

import numpy as np
import pytest
import torch

import depthpro_oracle as O
import reference_loader as RL
from depth_pro import weights


def test_split_merge_index_golden(golden_dir):
    gold = np.load(os.path.join(golden_dir, "split_merge_index.npz"))
    for name, size, ov in (("split_1536", 1536, 0.25), ("split_768", 768, 0.5)):
        x = torch.arange(size * size, dtype=torch.int32).reshape(1, 1, size, size)
        p = O.split(x, ov)
        assert np.array_equal(p[:, 0, 0, 0].numpy(), gold[name + "_corner"])
        assert np.array_equal(p[:, 0, -1, -1].numpy(), gold[name + "_last"])
        digest = np.frombuffer(hashlib.sha256(p.numpy().tobytes()).digest(), dtype=np.uint8)
        assert np.array_equal(digest, gold[name + "_sha256"])
        p2 = O.split(torch.stack([x[0], x[0] + size * size]), ov)
        assert np.array_equal(p2[:, 0, 0, 0].numpy(), gold[name + "_b2_corner"])
    for name, steps, pad in (("merge_5x5_pad3", 5, 3), ("merge_3x3_pad6", 3, 6)):
        n = steps * steps
        x = torch.arange(n * 576, dtype=torch.int32).reshape(n, 1, 24, 24)
        assert np.array_equal(O.merge(x, 1, pad)[0, 0].numpy(), gold[name])
        x2 = torch.arange(2 * n * 576, dtype=torch.int32).reshape(2 * n, 1, 24, 24)
        assert np.array_equal(O.merge(x2, 2, pad)[:, 0].numpy(), gold[name + "_b2"])


def test_merge_band_table():
    """5x5/pad3 -> 96 with bands [0,21),[21,39),...; 3x3/pad6 -> 48."""
    x = torch.arange(25 * 576, dtype=torch.int32).reshape(25, 1, 24, 24)
    m = O.merge(x, 1, 3)[0, 0]
    assert m.shape == (96, 96)
    rows = (m[:, 0] // 576 // 5).tolist()  # patch row j of every output row
    assert rows == [0] * 21 + [1] * 18 + [2] * 18 + [3] * 18 + [4] * 21
    x = torch.arange(9 * 576, dtype=torch.int32).reshape(9, 1, 24, 24)
    m = O.merge(x, 1, 6)[0, 0]
    assert m.shape == (48, 48)
    assert (m[:, 0] // 576 // 3).tolist() == [0] * 18 + [1] * 12 + [2] * 18


def test_pyramid_closed_form():
    """x1 = 2x2 box mean, x2 = mean of pixels (4i+1,4i+2)x(4j+1,4j+2)."""
    g = torch.Generator().manual_seed(0)
    x = torch.rand(1, 3, 64, 64, generator=g)
    _, x1, x2 = O.create_pyramid(x)
    box = 0.25 * (x[..., 0::2, 0::2] + x[..., 0::2, 1::2] + x[..., 1::2, 0::2] + x[..., 1::2, 1::2])
    assert torch.allclose(x1, box, atol=1e-6)
    mid = 0.25 * (x[..., 1::4, 1::4] + x[..., 1::4, 2::4] + x[..., 2::4, 1::4] + x[..., 2::4, 2::4])
    assert torch.allclose(x2, mid, atol=1e-6)


def test_depth_to_3d_golden(golden_dir):
    gold = np.load(os.path.join(golden_dir, "depth_to_3d.npz"))
    h, w = gold["depth"].shape
    pts, valid = O.depth_to_3d(gold["depth"], float(gold["f"]), w, h)
    assert pts.dtype == np.float64
    assert np.array_equal(valid, gold["valid"]) and np.array_equal(pts, gold["points"])
    # empty and all-invalid inputs
    pts0, valid0 = O.depth_to_3d(np.full((4, 5), np.nan, np.float32), 10.0, 5, 4)
    assert pts0.shape == (0, 3) and not valid0.any()


def test_full_infer_matches_reference_golden(golden_dir):
    """Oracle infer on the Config-1 input == the reference's recorded output (one CPU frame)."""
    torch.set_num_threads(os.cpu_count())
    gold = np.load(os.path.join(golden_dir, "reference_outputs.npz"))
    sd = weights.stress_init(1234)
    out = O.infer(sd, O.synthetic_image_1536(1))
    d = out["depth"][::16, ::16]
    g = torch.from_numpy(gold["depth_1536"])
    assert float(((d - g).abs() / g).max()) <= 2e-5
    assert abs(float(out["focallength_px"]) - float(gold["f_px_1536"])) <= 2e-5 * float(gold["f_px_1536"])


@pytest.mark.reference
@pytest.mark.skipif(not RL.available(), reason="/root/reference not present (GPU box)")
def test_oracle_vs_live_reference():
    """split / merge / one ViT-L forward of the oracle vs the reference's own code (+ timm shim)."""
    RL.load()
    enc = sys.modules["ref_depth_pro.network.encoder"].DepthProEncoder
    g = torch.Generator().manual_seed(1)
    x = torch.rand(2, 3, 1536, 1536, generator=g)
    assert torch.equal(enc.split(None, x, 0.25), O.split(x, 0.25))
    t = torch.rand(50, 16, 24, 24, generator=g)
    assert torch.equal(enc.merge(None, t, batch_size=2, padding=3), O.merge(t, 2, 3))
    vf = sys.modules["ref_depth_pro.network.vit_factory"]
    vit = vf.create_vit("dinov2l16_384").eval()
    sd = {"v." + k: v for k, v in vit.state_dict().items()}
    assert vit.pos_embed.shape == (1, 577, 1024) and vit.patch_embed.proj.weight.shape == (1024, 3, 16, 16)
    img = torch.rand(1, 3, 384, 384, generator=g) * 2 - 1
    with torch.no_grad():
        ref = vit(img)
        got, _ = O.vit_forward(sd, "v.", img)
    assert torch.equal(ref, got)


def test_ground_normalisation_oracle_matches_reference_outputs(golden_dir):
    """oracle restatement of img_to_normalized_pointcloud.py:880-1118 vs outputs of the reference's own functions
    (tests/golden/ground_normalize.npz, written by oracle/make_golden.py: rotation branch and |normal.y| > 0.99)."""
    g = np.load(os.path.join(golden_dir, "ground_normalize.npz"))
    for tag in ("tilt12", "tilt3"):
        pts, normal, d = g[tag + "_points"], g[tag + "_normal"], float(g[tag + "_d"])
        p2, n2, d2 = O.synthetic_room_points(20000, 3 if tag == "tilt12" else 4, 12.0 if tag == "tilt12" else 3.0)
        assert np.array_equal(p2, pts) and np.array_equal(n2, normal) and d2 == d      # seeded input is reproducible
        norm = O.normalize_point_cloud_to_ground(pts.astype(np.float64), normal, d)
        assert np.array_equal(norm.astype(np.float32), g[tag + "_normalized"])
        grid = O.grid_based_ground_adjustment(g[tag + "_normalized"].astype(np.float64), 20, 5)
        assert np.array_equal(grid.astype(np.float32), g[tag + "_grid"])
        assert (np.abs(normal[1]) > 0.99) == (tag == "tilt3")
        # the stages do something on this input: ground at y = 0, clamps hit, cells lowered
        assert (norm[:, 1] == 0).sum() > 10 and (norm[:, 1] == -0.1).sum() > 10
        assert (grid[:, 1] != g[tag + "_normalized"][:, 1].astype(np.float64)).sum() > 1000
    # degenerate inputs: fewer than 11 near-plane points -> no percentile shift; tiny clouds pass through the grid stage
    far = np.array([[0.0, 5.0, 1.0], [1.0, 6.0, 2.0], [2.0, 7.0, 3.0]])
    out = O.normalize_point_cloud_to_ground(far, np.array([0.0, 1.0, 0.0]), 0.0)
    assert np.array_equal(out, far)
    assert np.array_equal(O.grid_based_ground_adjustment(far, 20, 5), far)


@pytest.mark.reference
@pytest.mark.skipif(not RL.available(), reason="/root/reference not present (GPU box)")
def test_oracle_fov_head_without_encoder_vs_live_reference():
    """`fov_encoder_preset=None` (fov.py:55-56, 80-82): the oracle's conv-only FOV head against the reference's own
    FOVNetwork(num_features=256, fov_encoder=None), same state_dict keys (fov.head.{0,2,4,6}), bit for bit."""
    RL.load()
    FOV = sys.modules["ref_depth_pro.network.fov"].FOVNetwork
    net = FOV(num_features=256, fov_encoder=None).eval()
    want = weights.manifest("head")
    assert {"fov." + k: tuple(v.shape) for k, v in net.state_dict().items()} == {k: s for k, s in want.items() if k.startswith("fov.")}
    sd = {k: weights.stress_tensor(k, s, 3) for k, s in want.items() if k.startswith("fov.")}
    net.load_state_dict({k[4:]: v for k, v in sd.items()}, strict=True)
    g = torch.Generator().manual_seed(2)
    lowres = torch.randn(2, 256, 48, 48, generator=g)
    with torch.no_grad():
        ref = net(torch.zeros(2, 3, 1536, 1536), lowres)
    got = O.fov_forward(sd, torch.zeros(2, 3, 1536, 1536), lowres)
    assert ref.shape == (2, 1, 1, 1) and torch.equal(ref, got)
    assert O.fov_forward({"head.0.weight": torch.zeros(1)}, None, lowres) is None      # use_fov_head=False
    # the three manifests: default 1119 tensors, head-only drops the fov ViT, none drops every fov.* key
    assert len(weights.manifest()) == 1119
    assert not any(k.startswith("fov.") for k in weights.manifest(None))
    assert sum(k.startswith("fov.") for k in want) == 8


@pytest.mark.reference
@pytest.mark.skipif(not RL.available(), reason="/root/reference not present (GPU box)")
def test_vit_surgery_matches_the_reference_functions():
    """a13 (`resize_patch_embed` / `resize_vit`, vit.py:51-123): the product's state-dict form against the reference's
    own in-place surgery of a timm patch14 / 518 model, bit for bit, and the result loads into the engine's manifest."""
    from depth_pro import vit_surgery

    RL.load()
    import timm  # the shim (oracle/timm)

    vit_mod = sys.modules["ref_depth_pro.network.vit"]
    torch.manual_seed(3)
    raw_model = timm.create_model("vit_large_patch14_dinov2", pretrained=False, dynamic_img_size=True)
    raw = {k: v.clone() for k, v in raw_model.state_dict().items()}
    assert raw["patch_embed.proj.weight"].shape == (1024, 3, 14, 14) and raw["pos_embed"].shape == (1, 1370, 1024)
    raw_model.patch_size = raw_model.patch_embed.patch_size
    ref_model = vit_mod.resize_vit(vit_mod.resize_patch_embed(raw_model, new_patch_size=(16, 16)), img_size=(384, 384))
    ref = ref_model.state_dict()
    got = vit_surgery.convert_timm_vit_state_dict(raw, prefix="encoder.patch_encoder.")
    want_shapes = {k: s for k, s in weights.manifest().items() if k.startswith("encoder.patch_encoder.")}
    assert {k: tuple(v.shape) for k, v in got.items()} == want_shapes
    for k, v in ref.items():
        assert torch.equal(got["encoder.patch_encoder." + k], v), k
    assert got["encoder.patch_encoder.pos_embed"].shape == (1, 577, 1024)
