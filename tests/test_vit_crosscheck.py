"""Independent pin of the third-party ViT boundary.

The reference builds its three encoders with `timm.create_model("vit_large_patch14_dinov2", ...)`
(/root/reference/src/depth_pro/network/vit_factory.py:97-110) and calls `forward_features` on them
(/root/reference/src/depth_pro/network/vit.py:33; hooks on blocks 5 and 11 at encoder.py:133-144).  timm is not in
this image, so `oracle/timm/` and `oracle.vit_forward` RESTATE it — two restatements by one author.  This test checks
them against an implementation the builder did not write: `transformers.models.dinov2.Dinov2Model` (HF port of the
same DINOv2 ViT-L: cls token + 577-token position embedding, pre-norm blocks, fused-qkv split in q/k/v, exact GELU,
LayerNorm eps 1e-6, LayerScale after attention and after the MLP, final LayerNorm), loaded with the SAME seeded
weights through a key map.
"""

import pytest
import torch

import depthpro_oracle as O
from depth_pro import weights

transformers = pytest.importorskip("transformers")


def _hf_dinov2_from_timm_keys(sd, prefix, depth=24):
    from transformers import Dinov2Config, Dinov2Model

    cfg = Dinov2Config(hidden_size=1024, num_hidden_layers=depth, num_attention_heads=16, mlp_ratio=4,
                       image_size=384, patch_size=16, layer_norm_eps=1e-6, hidden_act="gelu", qkv_bias=True,
                       use_swiglu_ffn=False, layerscale_value=1.0, hidden_dropout_prob=0.0,
                       attention_probs_dropout_prob=0.0, drop_path_rate=0.0)
    model = Dinov2Model(cfg).eval()
    w = lambda k: sd[prefix + k]
    hf = {"embeddings.cls_token": w("cls_token"), "embeddings.position_embeddings": w("pos_embed"),
          "embeddings.mask_token": torch.zeros(1, 1024),
          "embeddings.patch_embeddings.projection.weight": w("patch_embed.proj.weight"),
          "embeddings.patch_embeddings.projection.bias": w("patch_embed.proj.bias"),
          "layernorm.weight": w("norm.weight"), "layernorm.bias": w("norm.bias")}
    for i in range(depth):
        b, o = f"blocks.{i}.", f"encoder.layer.{i}."
        qw, kw, vw = w(b + "attn.qkv.weight").split(1024, dim=0)   # timm: rows [q | k | v], each (heads*64, 1024)
        qb, kb, vb = w(b + "attn.qkv.bias").split(1024, dim=0)
        hf.update({o + "attention.attention.query.weight": qw, o + "attention.attention.query.bias": qb,
                   o + "attention.attention.key.weight": kw, o + "attention.attention.key.bias": kb,
                   o + "attention.attention.value.weight": vw, o + "attention.attention.value.bias": vb,
                   o + "attention.output.dense.weight": w(b + "attn.proj.weight"),
                   o + "attention.output.dense.bias": w(b + "attn.proj.bias"),
                   o + "layer_scale1.lambda1": w(b + "ls1.gamma"), o + "layer_scale2.lambda1": w(b + "ls2.gamma"),
                   o + "norm1.weight": w(b + "norm1.weight"), o + "norm1.bias": w(b + "norm1.bias"),
                   o + "norm2.weight": w(b + "norm2.weight"), o + "norm2.bias": w(b + "norm2.bias"),
                   o + "mlp.fc1.weight": w(b + "mlp.fc1.weight"), o + "mlp.fc1.bias": w(b + "mlp.fc1.bias"),
                   o + "mlp.fc2.weight": w(b + "mlp.fc2.weight"), o + "mlp.fc2.bias": w(b + "mlp.fc2.bias")})
    model.load_state_dict({k: v.clone() for k, v in hf.items()}, strict=True)
    return model


@pytest.fixture(scope="module")
def sd():
    return weights.stress_init(1234)


@pytest.mark.parametrize("prefix", ["encoder.patch_encoder.", "fov.encoder.0."])
def test_oracle_vit_matches_transformers_dinov2(sd, prefix):
    """All 577x1024 normed outputs and the block-5 / block-11 hook activations of `oracle.vit_forward` equal the HF
    Dinov2Model's to <= 1e-5 (relative to the tensor's absmax), for two encoders' (different) seeded weights."""
    x = O.synthetic_image_1536(3)[None, :, 100:484, 200:584].contiguous()
    x = torch.cat([x, x.flip(-1) * 0.5], dim=0)                      # two different 384^2 patches
    with torch.no_grad():
        out, hooks = O.vit_forward(sd, prefix, x, hook_ids=(5, 11))
        hf = _hf_dinov2_from_timm_keys(sd, prefix)(pixel_values=x, output_hidden_states=True)
    assert out.shape == (2, 577, 1024) == hf.last_hidden_state.shape
    rel = lambda a, b: float((a - b).abs().max() / b.abs().max())
    assert rel(out, hf.last_hidden_state) <= 1e-5
    assert torch.allclose(out, hf.last_hidden_state, rtol=1e-5, atol=1e-5 * float(out.abs().max()))
    # hidden_states[0] = embeddings, hidden_states[i + 1] = output of block i (before the final norm)
    assert rel(hooks[0], hf.hidden_states[6]) <= 1e-5
    assert rel(hooks[1], hf.hidden_states[12]) <= 1e-5
    # the cross-check has teeth: it is sensitive to the token order and to the LayerScale placement
    assert rel(out[:, 1:], hf.last_hidden_state[:, :-1]) > 1e-2


@pytest.mark.reference
def test_reference_create_vit_on_shim_matches_transformers_dinov2(sd):
    """The UNMODIFIED reference `create_vit("dinov2l16_384")` (vit_factory.py:55-124: timm.create_model + the
    patch-embed / pos-embed surgery of vit.py:51-123) running on `oracle/timm`, against HF Dinov2Model with the same
    (post-surgery) tensors.  Build container only: /root/reference does not exist on the GPU box."""
    import sys

    import reference_loader as RL

    if not RL.available():
        pytest.skip("/root/reference not present (GPU box)")
    RL.load()
    vit = sys.modules["ref_depth_pro.network.vit_factory"].create_vit("dinov2l16_384").eval()
    prefix = "encoder.image_encoder."
    vit.load_state_dict({k[len(prefix):]: v for k, v in sd.items() if k.startswith(prefix)}, strict=True)
    x = O.synthetic_image_1536(2)[None, :, :384, :384].contiguous()
    with torch.no_grad():
        a = vit(x)
        b = _hf_dinov2_from_timm_keys(sd, prefix)(pixel_values=x).last_hidden_state
    assert a.shape == b.shape == (1, 577, 1024)
    assert float((a - b).abs().max() / b.abs().max()) <= 1e-5
