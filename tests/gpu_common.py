"""Helpers shared by the -m gpu tests: one engine handle per precision, ctypes call sugar."""

import ctypes

import torch

from depth_pro import _capi

_engines = {}


def lib(flavour="bf16"):
    return _capi.load(flavour)


def engine(prec=_capi.PREC_FP32, flavour="bf16"):
    """A bare engine (no weights) — enough for the kernel-level entry points."""
    if (prec, flavour) not in _engines:
        h = ctypes.c_void_p()
        _capi.check(lib(flavour).dp_engine_create(0, prec, 1, ctypes.byref(h)))
        _engines[(prec, flavour)] = h
    return _engines[(prec, flavour)]


def stream():
    return torch.cuda.current_stream().cuda_stream


def relerr(a, b):
    return float((a.double() - b.double()).abs().max() / b.double().abs().max().clamp_min(1e-30))
