"""Helpers shared by the -m gpu tests: one engine handle per precision, ctypes call sugar."""

import ctypes

import torch

from depth_pro import _capi

_engines = {}


def lib():
    return _capi.load()


def engine(prec=_capi.PREC_FP32):
    """A bare engine (no weights) — enough for the kernel-level entry points."""
    if prec not in _engines:
        h = ctypes.c_void_p()
        _capi.check(lib().dp_engine_create(0, prec, 1, ctypes.byref(h)))
        _engines[prec] = h
    return _engines[prec]


def stream():
    return torch.cuda.current_stream().cuda_stream


def relerr(a, b):
    return float((a.double() - b.double()).abs().max() / b.double().abs().max().clamp_min(1e-30))
