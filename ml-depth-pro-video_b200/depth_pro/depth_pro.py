"""Drop-in replacement for the reference's ``depth_pro.depth_pro`` module.

Same public surface as ``/root/reference/src/depth_pro/depth_pro.py``:
``DepthProConfig`` (:26-36), ``DEFAULT_MONODEPTH_CONFIG_DICT`` (:39-46),
``create_model_and_transforms(config, device, precision) -> (model, transform)`` (:72-151) and
``DepthPro`` with ``.img_size`` (:213-216), ``.forward(x)`` (:218-241), ``.infer(x, f_px,
interpolation_mode)`` (:243-298), ``.eval()``, ``.state_dict()`` / ``.load_state_dict()`` with the
reference's key names.  All compute runs in the sm_100a engine behind the C-ABI
(``include/depthpro_b200.h``); there is no PyTorch / CPU fallback.
"""

from __future__ import annotations

import ctypes
import logging
from dataclasses import dataclass
from typing import Mapping, Optional, Tuple, Union

import numpy as np
import torch
from torch import nn

from . import _capi, weights

LOGGER = logging.getLogger(__name__)

ViTPreset = str
_PRESETS = ("dinov2l16_384",)
IMG_SIZE = 1536


@dataclass
class DepthProConfig:
    """Configuration for DepthPro (same fields as the reference's dataclass)."""

    patch_encoder_preset: ViTPreset
    image_encoder_preset: ViTPreset
    decoder_features: int

    checkpoint_uri: Optional[str] = None
    fov_encoder_preset: Optional[ViTPreset] = None
    use_fov_head: bool = True


DEFAULT_MONODEPTH_CONFIG_DICT = DepthProConfig(
    patch_encoder_preset="dinov2l16_384",
    image_encoder_preset="dinov2l16_384",
    checkpoint_uri="./checkpoints/depth_pro.pt",
    decoder_features=256,
    use_fov_head=True,
    fov_encoder_preset="dinov2l16_384",
)


class _Node(nn.Module):
    """Parameter container; the tree of _Nodes reproduces the reference's module names."""


def _precision_code(precision: torch.dtype) -> int:
    if precision == torch.float32:
        return _capi.PREC_FP32
    if precision in (torch.bfloat16, torch.float16):
        return _capi.PREC_BF16   # "the 16-bit mode" of the library flavour `_flavour` selects
    raise ValueError(f"unsupported precision {precision}; use torch.float32, torch.float16 or torch.bfloat16")


def _flavour(precision: torch.dtype) -> str:
    """Which build of the engine serves this precision: torch.half (the reference's `model.half()`, depth_pro.py:122-123)
    runs on libdepthpro_b200_fp16.so (16-bit storage = IEEE half), everything else on libdepthpro_b200.so (bfloat16 /
    fp32).  Both keep fp32 accumulation, an fp32 ViT residual stream and fp32 statistics."""
    return "fp16" if precision == torch.float16 else "bf16"


class DepthPro(nn.Module):
    """Depth Pro network backed by the B200 engine."""

    def __init__(self, device: torch.device, precision: torch.dtype = torch.float32, max_batch: int = 1,
                 use_fov_head: bool = True, fov_encoder: bool = True):
        """``use_fov_head`` / ``fov_encoder`` mirror the reference constructor (depth_pro.py:153-211): without the head
        ``forward`` returns ``fov_deg=None`` and ``infer`` needs ``f_px``; ``fov_encoder=False`` is the head without
        its own ViT (``DepthProConfig.fov_encoder_preset=None``)."""
        super().__init__()
        device = torch.device(device)
        if device.type != "cuda":
            raise RuntimeError(
                f"depth_pro (B200 engine) needs a CUDA device, got '{device}': there is no CPU fallback")
        if not torch.cuda.is_available():
            raise RuntimeError("depth_pro (B200 engine): no CUDA device is visible; there is no CPU fallback")
        self._device = torch.device("cuda", device.index if device.index is not None else torch.cuda.current_device())
        self._precision = precision
        self._prec_code = _precision_code(precision)
        self._lib_flavour = _flavour(precision)
        self._max_batch = int(max_batch)
        self._engine = None
        self._dirty = True
        self._fov = None if not use_fov_head else ("encoder" if fov_encoder else "head")
        for name, shape in weights.manifest(self._fov).items():
            node: nn.Module = self
            parts = name.split(".")
            for p in parts[:-1]:
                if not hasattr(node, p):
                    node.add_module(p, _Node())
                node = getattr(node, p)
            node.register_parameter(
                parts[-1], nn.Parameter(torch.empty(shape, dtype=torch.float32, device=self._device),
                                        requires_grad=False))
        self.register_load_state_dict_post_hook(lambda module, incompatible: module._mark_dirty())

    # ------------------------------------------------------------------ weights / engine
    def _mark_dirty(self):
        self._dirty = True

    def refresh_weights(self) -> "DepthPro":
        """Re-upload every parameter to the engine before the next call.  ``load_state_dict``, ``init_weights`` and
        module-level conversions (``.to()``, ``.half()``, ``.float()``) are tracked automatically; IN-PLACE edits of a
        parameter (``p.data.copy_(...)``, ``p.mul_(...)``) are not -- call this after them, otherwise the engine keeps
        computing with its packed copy of the old values."""
        self._dirty = True
        return self

    def _apply(self, fn, recurse=True):
        out = super()._apply(fn, recurse)
        self._dirty = True
        return out

    def init_weights(self, recipe: str = "stress", seed: int = 1234) -> "DepthPro":
        """Seeded random init (no checkpoint offline): 'stress' = recipe B, 'reference' = the
        reference's own (degenerate) distributions.  See weights.py."""
        fn = {"stress": weights.stress_tensor, "reference": weights.reference_like_tensor}[recipe]
        with torch.no_grad():
            for name, p in self.named_parameters():
                p.copy_(fn(name, tuple(p.shape), seed))
        self._dirty = True
        return self

    def _ensure_engine(self, batch: int):
        lib = _capi.load(self._lib_flavour)
        if self._engine is not None and batch > self._max_batch:
            _capi.check(lib.dp_engine_destroy(self._engine))
            self._engine = None
        if self._engine is None:
            self._max_batch = max(self._max_batch, batch)
            h = ctypes.c_void_p()
            fov_mode = {None: _capi.FOV_NONE, "head": _capi.FOV_HEAD_ONLY, "encoder": _capi.FOV_ENCODER}[self._fov]
            _capi.check(lib.dp_engine_create_ex(self._device.index, self._prec_code, self._max_batch, fov_mode,
                                                ctypes.byref(h)))
            self._engine = h
            self._dirty = True
        if self._dirty:
            torch.cuda.synchronize(self._device)
            for name, p in self.named_parameters():
                t = p.detach()
                if t.dtype != torch.float32 or not t.is_contiguous():
                    t = t.float().contiguous()
                shape = (ctypes.c_int64 * t.dim())(*t.shape)
                _capi.check(lib.dp_engine_set_weight(self._engine, name.encode(), t.data_ptr(), shape, t.dim(),
                                                     1 if t.is_cuda else 0))
            _capi.check(lib.dp_engine_finalize(self._engine))
            self._dirty = False
        return lib

    def __del__(self):
        try:
            if getattr(self, "_engine", None) is not None:
                _capi.load(self._lib_flavour).dp_engine_destroy(self._engine)
                self._engine = None
        except Exception:
            pass

    def _stream(self) -> int:
        return torch.cuda.current_stream(self._device).cuda_stream

    # ------------------------------------------------------------------ reference API
    @property
    def img_size(self) -> int:
        """Internal image size of the network (depth_pro.py:213-216)."""
        return IMG_SIZE

    def forward(self, x: torch.Tensor) -> Tuple[torch.Tensor, Optional[torch.Tensor]]:
        """(B,3,1536,1536) -> (canonical inverse depth (B,1,1536,1536), fov_deg (B,1,1,1))."""
        _, _, H, W = x.shape
        assert H == self.img_size and W == self.img_size
        x = x.to(device=self._device, dtype=torch.float32).contiguous()
        B = x.shape[0]
        lib = self._ensure_engine(B)
        canon = torch.empty((B, 1, IMG_SIZE, IMG_SIZE), dtype=torch.float32, device=self._device)
        fov = torch.empty((B, 1, 1, 1), dtype=torch.float32, device=self._device) if self._fov is not None else None
        with torch.cuda.device(self._device):
            _capi.check(lib.dp_forward(self._engine, x.data_ptr(), B, canon.data_ptr(), _capi.ptr(fov), self._stream()))
        return canon, fov   # fov is None without the FOV head, like depth_pro.py:236-241

    @torch.no_grad()
    def infer(self, x: torch.Tensor, f_px: Optional[Union[float, torch.Tensor]] = None,
              interpolation_mode="bilinear") -> Mapping[str, torch.Tensor]:
        """Depth [m] and focal length [px] for an image (depth_pro.py:243-298).

        ``x`` is the reference's transformed tensor (float CHW / BCHW in [-1,1]); additionally a
        uint8 HWC / BHWC tensor or ndarray straight from ``load_rgb`` is accepted, in which case
        ToTensor + Normalize are fused into the resize kernel.
        """
        if interpolation_mode not in _capi.INTERP:
            # F.interpolate(..., mode=m, align_corners=False) on a 4-D tensor accepts "bilinear" and "bicubic" only; the
            # reference raises this ValueError for every other mode ("nearest", "area", "nearest-exact", ...)
            raise ValueError("align_corners option can only be set with the interpolating modes: "
                             "linear | bilinear | bicubic | trilinear")
        if f_px is None and self._fov is None:
            # depth_pro.py:282-283 dereferences fov_deg=None here
            raise AttributeError("'NoneType' object has no attribute 'to' (this model has no FOV head: pass f_px)")
        if isinstance(x, np.ndarray):
            x = torch.from_numpy(np.ascontiguousarray(x))
        u8 = x.dtype == torch.uint8
        if len(x.shape) == 3:
            x = x.unsqueeze(0)
        if u8:
            B, H, W, ch = x.shape
            assert ch == 3, "uint8 input must be HWC with 3 channels"
            x = x.to(self._device).contiguous()
            fmt = _capi.SRC_U8_HWC
        else:
            B, ch, H, W = x.shape
            assert ch == 3
            x = x.to(device=self._device, dtype=torch.float32).contiguous()
            fmt = _capi.SRC_F32_CHW
        lib = self._ensure_engine(B)
        depth = torch.empty((B, H, W), dtype=torch.float32, device=self._device)
        f_out = torch.empty((B,), dtype=torch.float32, device=self._device)
        f_host = None
        if f_px is not None:
            f_t = torch.as_tensor(f_px).detach().to("cpu", torch.float32).reshape(-1)
            if f_t.numel() == 1:
                f_t = f_t.expand(B)
            assert f_t.numel() == B, "f_px must be a scalar or one value per image"
            f_host = f_t.contiguous()
        with torch.cuda.device(self._device):
            _capi.check(lib.dp_infer_ex(self._engine, x.data_ptr(), B, H, W, fmt, _capi.INTERP[interpolation_mode],
                                        None if f_host is None else f_host.data_ptr(),
                                        depth.data_ptr(), f_out.data_ptr(), self._stream()))
        if f_px is None:
            focal = f_out.squeeze()
        else:
            focal = f_px.squeeze() if torch.is_tensor(f_px) else torch.as_tensor(np.asarray(f_px)).squeeze()
        return {"depth": depth.squeeze(), "focallength_px": focal}

    # ------------------------------------------------------------------ engine-level extras
    def tap(self, stage: str) -> torch.Tensor:
        """Stage tensor of the last forward as float32 in the reference's NCHW layout (tests)."""
        lib = self._ensure_engine(1)
        cap = 768 * 768 * 256 * max(1, self._max_batch)
        out = torch.empty(cap, dtype=torch.float32, device=self._device)
        n = ctypes.c_int64()
        _capi.check(lib.dp_tap(self._engine, stage.encode(), out.data_ptr(), cap, ctypes.byref(n), self._stream()))
        return out[: n.value]

    def launch_count(self) -> int:
        return int(_capi.load(self._lib_flavour).dp_launch_count(self._engine))


def create_backbone_model(preset: ViTPreset):
    """Kept for signature compatibility (depth_pro.py:49-69): validates the preset only."""
    if preset not in _PRESETS:
        raise KeyError(f"Preset {preset} not found.")
    return None, preset


def create_model_and_transforms(
    config: DepthProConfig = DEFAULT_MONODEPTH_CONFIG_DICT,
    device: torch.device = torch.device("cpu"),
    precision: torch.dtype = torch.float32,
):
    """Create a DepthPro model and load weights from ``config.checkpoint_uri`` (depth_pro.py:72-151).

    ``device`` must be a CUDA device (B200).  With ``checkpoint_uri=None`` the parameters get the
    reference's (degenerate) random-init distributions; call ``model.init_weights('stress')`` for
    the conditioned recipe used by the parity tests and the benchmark.
    """
    from torchvision.transforms import Compose, ConvertImageDtype, Lambda, Normalize, ToTensor

    for preset in (config.patch_encoder_preset, config.image_encoder_preset):
        create_backbone_model(preset)
    fov_encoder = config.use_fov_head and config.fov_encoder_preset is not None     # depth_pro.py:100-102
    if fov_encoder:
        create_backbone_model(config.fov_encoder_preset)
    if config.decoder_features != 256:
        # every tensor-core tile shape, the composed head and the workspace plan are built for dim_decoder = 256 (the only
        # value a published checkpoint has); rejected loudly rather than run on a silent slow path
        raise NotImplementedError("the B200 engine implements decoder_features=256")

    model = DepthPro(device=device, precision=precision, use_fov_head=config.use_fov_head, fov_encoder=fov_encoder)
    transform = Compose([
        ToTensor(),
        Lambda(lambda x: x.to(device)),
        Normalize([0.5, 0.5, 0.5], [0.5, 0.5, 0.5]),
        ConvertImageDtype(torch.float32),  # the engine takes fp32 input in every precision mode
    ])

    if config.checkpoint_uri is not None:
        state_dict = torch.load(config.checkpoint_uri, map_location="cpu")
        missing_keys, unexpected_keys = model.load_state_dict(state_dict=state_dict, strict=True)
        if len(unexpected_keys) != 0:
            raise KeyError(f"Found unexpected keys when loading monodepth: {unexpected_keys}")
        missing_keys = [key for key in missing_keys if "fc_norm" not in key]
        if len(missing_keys) != 0:
            raise KeyError(f"Keys are missing when loading monodepth: {missing_keys}")
    else:
        model.init_weights("reference", seed=0)
    return model, transform


__all__ = ["DepthPro", "DepthProConfig", "DEFAULT_MONODEPTH_CONFIG_DICT", "create_model_and_transforms",
           "create_backbone_model"]
