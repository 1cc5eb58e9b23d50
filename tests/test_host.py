"""CPU tests of the host side: manifest, seeded init, C-ABI symbol table, error behaviour, IO."""

import ctypes
import json
import os
import re

import numpy as np
import pytest
import torch

import depth_pro
from depth_pro import _capi, weights

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_manifest_matches_reference_state_dict(golden_dir):
    gold = json.load(open(os.path.join(golden_dir, "state_dict_manifest.json")))
    m = weights.manifest()
    assert list(gold) == list(m)
    assert all(tuple(gold[k]) == tuple(m[k]) for k in m)
    assert sum(int(np.prod(s)) for s in m.values()) == 951_991_330


def test_seeded_init_is_deterministic_and_order_free():
    a = weights.stress_tensor("head.0.weight", (128, 256, 3, 3), 1234)
    b = weights.stress_tensor("head.0.weight", (128, 256, 3, 3), 1234)
    c = weights.stress_tensor("head.0.weight", (128, 256, 3, 3), 1235)
    assert torch.equal(a, b) and not torch.equal(a, c)
    assert abs(float(a.std()) - 1 / (9 * 256) ** 0.5) < 1e-3
    assert float(weights.stress_tensor("head.4.bias", (1,), 1)) == 2.0
    assert float(weights.stress_tensor("fov.head.4.bias", (1,), 1)) == 60.0
    g = weights.stress_tensor("encoder.patch_encoder.blocks.3.ls1.gamma", (1024,), 7)
    assert 0.05 <= float(g.min()) and float(g.max()) <= 0.3
    # ConvTranspose weights are (Cin, Cout, 2, 2): std 1/sqrt(Cin)
    t = weights.stress_tensor("encoder.upsample0.1.weight", (512, 512, 2, 2), 3)
    assert abs(float(t.std()) - 512 ** -0.5) < 1e-3


def test_capi_exports_every_declared_symbol():
    header = open(os.path.join(ROOT, "include", "depthpro_b200.h")).read()
    declared = set(re.findall(r"\b(dp_[a-z0-9_]+)\s*\(", header))
    assert declared == set(_capi.SIGNATURES), declared ^ set(_capi.SIGNATURES)
    for flavour in ("bf16", "fp16"):   # both builds of the library (16-bit storage = bfloat16 / IEEE half)
        lib = _capi.load(flavour)      # loads without a GPU; no compute call is made here
        for name in declared:
            assert hasattr(lib, name), (flavour, name)
        assert lib.dp_version() >= 200 and lib.dp_act_dtype().decode() == flavour


def test_engine_create_fails_loudly_without_gpu():
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    h = ctypes.c_void_p()
    rc = _capi.load().dp_engine_create(0, 0, 1, ctypes.byref(h))
    assert rc != 0 and len(_capi.load().dp_last_error()) > 0
    with pytest.raises(RuntimeError):
        depth_pro.create_model_and_transforms(device=torch.device("cpu"))
    with pytest.raises(RuntimeError):
        depth_pro.DepthPro(device=torch.device("cuda:0"))


def test_config_and_errors():
    cfg = depth_pro.DEFAULT_MONODEPTH_CONFIG_DICT
    assert cfg.patch_encoder_preset == "dinov2l16_384" and cfg.decoder_features == 256
    assert cfg.checkpoint_uri == "./checkpoints/depth_pro.pt" and cfg.use_fov_head
    with pytest.raises(KeyError):
        depth_pro.depth_pro.create_backbone_model("unknown_preset")
    with pytest.raises(ValueError):
        depth_pro.depth_pro._precision_code(torch.int8)


def test_load_rgb_contract(tmp_path):
    from PIL import Image

    arr = (np.random.default_rng(0).random((20, 30, 4)) * 255).astype(np.uint8)
    p = tmp_path / "a.png"
    Image.fromarray(arr, "RGBA").save(p)
    img, icc, f_px = depth_pro.load_rgb(p)
    assert img.dtype == np.uint8 and img.shape == (20, 30, 3) and f_px is None
    assert np.array_equal(img, arr[:, :, :3])
    gray = tmp_path / "g.png"
    Image.fromarray(arr[:, :, 0], "L").save(gray)
    assert depth_pro.load_rgb(gray)[0].shape == (20, 30, 3)
    assert abs(depth_pro.utils.fpx_from_f35(36, 24, 50) - 50.0) < 1e-9


def test_synthetic_generators_match_oracle_copy():
    import depthpro_oracle as O
    from depth_pro import synthetic

    assert torch.equal(synthetic.synthetic_image_1536(3), O.synthetic_image_1536(3))
    assert np.array_equal(synthetic.synthetic_frame_u8(2, 90, 160), O.synthetic_frame_u8(2, 90, 160))


def test_reference_call_sites_are_covered_by_the_drop_in_package():
    """Every `depth_pro.<name>` the reference's scripts touch, and every keyword they pass to
    create_model_and_transforms / infer, exists with the same spelling in the drop-in package (skipped where
    /root/reference is not mounted, e.g. on the GPU box)."""
    import ast
    import inspect

    ref = "/root/reference"
    if not os.path.isdir(ref):
        pytest.skip("reference checkout not present")
    scripts = ["generate_depth_maps.py", "img_to_normalized_pointcloud.py", "pointcloud_cleaner.py", "pointcloud_to_mesh.py",
               "src/depth_pro/cli/run.py"]
    attrs, create_kw, infer_kw = set(), set(), set()
    for rel in scripts:
        tree = ast.parse(open(os.path.join(ref, rel)).read())
        for node in ast.walk(tree):
            if isinstance(node, ast.Attribute) and isinstance(node.value, ast.Name) and node.value.id == "depth_pro":
                attrs.add(node.attr)
            if isinstance(node, ast.Call) and isinstance(node.func, ast.Attribute):
                if node.func.attr == "create_model_and_transforms":
                    create_kw |= {k.arg for k in node.keywords}
                if node.func.attr == "infer":
                    infer_kw |= {k.arg for k in node.keywords}
    assert {"create_model_and_transforms", "load_rgb"} <= attrs
    for a in attrs:
        assert hasattr(depth_pro, a), f"reference scripts use depth_pro.{a}"
    assert create_kw <= set(inspect.signature(depth_pro.create_model_and_transforms).parameters), create_kw
    assert infer_kw <= set(inspect.signature(depth_pro.DepthPro.infer).parameters), infer_kw
    # the reference's signatures, parameter for parameter (depth_pro.py:72-76, 243-249)
    assert list(inspect.signature(depth_pro.create_model_and_transforms).parameters) == ["config", "device", "precision"]
    assert list(inspect.signature(depth_pro.DepthPro.infer).parameters) == ["self", "x", "f_px", "interpolation_mode"]
    assert list(inspect.signature(depth_pro.load_rgb).parameters) == ["path", "auto_rotate", "remove_alpha"]


def test_environment_switches_are_documented():
    """Every DEPTHPRO_* switch the sources read is listed in INTEGRATION.md §E / README.md, and every documented one
    is read somewhere (documentation that names a switch the code no longer honours is worse than none)."""
    import glob
    import re

    doc = open(os.path.join(ROOT, "INTEGRATION.md")).read() + open(os.path.join(ROOT, "README.md")).read()
    documented = set(re.findall(r"DEPTHPRO_[A-Z0-9_]+", doc))
    pkg = os.path.join(ROOT, "ml-depth-pro-video_b200")
    files = (glob.glob(os.path.join(pkg, "csrc", "*.cu")) + glob.glob(os.path.join(pkg, "csrc", "*.cuh"))
             + glob.glob(os.path.join(pkg, "depth_pro", "*.py")) + [os.path.join(ROOT, "bench.py")]
             + glob.glob(os.path.join(ROOT, "tests", "test_gpu_*.py")))
    src = "".join(open(f).read() for f in files)
    used = set(re.findall(r'getenv\("(DEPTHPRO_[A-Z0-9_]+)"\)', src))
    used |= set(re.findall(r'environ(?:\.get)?[\(\[]"(DEPTHPRO_[A-Z0-9_]+)"', src))
    assert used, "no switches found: the patterns above no longer match the sources"
    assert used - documented == set(), f"undocumented switches: {sorted(used - documented)}"
    assert documented - used == set(), f"documented but unused switches: {sorted(documented - used)}"


def test_weight_composition_algebra_out_conv_into_head0():
    """The algebra behind `compose_1x1_conv3x3` + `GemmOp::border_cb` (csrc/kernels.cu, gemm_tc.cu), restated in torch on
    the CPU: a 1x1 conv with bias (decoder.fusions.0.out_conv, decoder.py:178) followed by a zero-padded 3x3 conv with
    bias (head.0, depth_pro.py:183-185) equals ONE 3x3 conv with weights w3 . wo, the interior bias b3 + sum over all
    nine taps of w3[:, :, tap] . bo, minus -- on the outermost pixel ring only -- the share of every tap that falls
    into the padding (the 1x1's bias does not exist there)."""
    g = torch.Generator().manual_seed(3)
    C, O, H, W = 8, 6, 7, 9
    x = torch.randn(1, C, H, W, generator=g, dtype=torch.float64)
    wo, bo = torch.randn(C, C, 1, 1, generator=g, dtype=torch.float64), torch.randn(C, generator=g, dtype=torch.float64)
    w3, b3 = torch.randn(O, C, 3, 3, generator=g, dtype=torch.float64), torch.randn(O, generator=g, dtype=torch.float64)
    ref = torch.nn.functional.conv2d(torch.nn.functional.conv2d(x, wo, bo), w3, b3, padding=1)
    wc = torch.einsum("ockl,ci->oikl", w3, wo[:, :, 0, 0])            # [O][i][ky][kx]
    cb = torch.einsum("ockl,c->klo", w3, bo).reshape(9, O)              # per-tap share of bo
    out = torch.nn.functional.conv2d(x, wc, b3 + cb.sum(0), padding=1)
    for y in range(H):
        for xx in range(W):
            for ky in range(3):
                for kx in range(3):
                    if not (0 <= y + ky - 1 < H and 0 <= xx + kx - 1 < W):
                        out[0, :, y, xx] -= cb[ky * 3 + kx]
    assert torch.allclose(out, ref, rtol=1e-12, atol=1e-12)
    interior = torch.nn.functional.conv2d(x, wc, b3 + cb.sum(0), padding=1)
    assert torch.allclose(interior[..., 1:-1, 1:-1], ref[..., 1:-1, 1:-1], rtol=1e-12, atol=1e-12)
    assert not torch.allclose(interior, ref)                           # ... and the ring really needs the correction


def test_pair_residual_stream_arithmetic():
    """The (hi, lo) form of the ViT residual stream (csrc/common.cuh GemmOp::ln_xlo), restated in torch on the CPU:
    hi = bf16(x), lo = bf16(x - hi) carries x to 2^-16 of its value (x - hi is exact in fp32: at most 16 significant bits
    are left, bf16 keeps 8 of them), hi alone is exactly the operand the next GEMM would have read from a separate copy,
    and 48 consecutive updates (a ViT-L's 24 proj + 24 fc2) stay within 48 * 2^-17 of the fp32 stream."""
    g = torch.Generator().manual_seed(4)
    x = torch.randn(4096, generator=g) * torch.logspace(-3, 3, 4096)

    def split(v):
        hi = v.bfloat16()
        return hi, (v - hi.float()).bfloat16()

    hi, lo = split(x)
    assert torch.equal(hi, x.bfloat16())
    back = hi.float() + lo.float()
    assert float(((back - x).abs() / x.abs()).max()) <= 2.0 ** -16
    x32 = x.clone()
    scale = x.abs()                                    # largest magnitude the element has had so far
    for i in range(48):
        d = torch.randn(4096, generator=g) * 0.1 * x.abs()
        x32 = x32 + d
        hi, lo = split(hi.float() + lo.float() + d)
        scale = torch.maximum(scale, x32.abs())
    # every split rounds by at most 2^-17 of the value AT THAT TIME (half an ulp of lo's 8-bit mantissa)
    assert float((((hi.float() + lo.float()) - x32).abs() / scale).max()) <= 48 * 2.0 ** -17
