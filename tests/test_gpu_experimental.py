"""-m gpu, and only with DEPTHPRO_TEST_EXPERIMENTAL=1: checks for the opt-in code paths that were written in the last,
GPU-less hours of round 1 and have NOT been run on a GPU yet: attention variants 12-14 (P through TMEM,
second Q buffer) and the four-pixels-per-thread metric-depth epilogue.  They are skipped in the normal suite so that an
unvalidated experiment can never take the `-x` run down; once a variant has passed here on a B200 its case moves into
the regular test files.

    DEPTHPRO_TEST_EXPERIMENTAL=1 python -m pytest tests/test_gpu_experimental.py -m gpu -q
"""

import ctypes
import os

import pytest
import torch
import torch.nn.functional as F

from gpu_common import engine, lib, relerr, stream
from depth_pro import _capi

pytestmark = [pytest.mark.gpu,
              pytest.mark.skipif(os.environ.get("DEPTHPRO_TEST_EXPERIMENTAL") != "1",
                                 reason="experimental paths: set DEPTHPRO_TEST_EXPERIMENTAL=1")]
DEV = "cuda:0"


def _kb(kind, M, N, K, iters=1):
    ms = ctypes.c_float()
    _capi.check(lib().dp_kernel_bench(engine(), kind, M, N, K, iters, ctypes.byref(ms)))
    return ms.value


@pytest.mark.parametrize("H,W", [(1536, 1536), (1080, 1920), (2160, 3840), (333, 2001), (7, 5), (1, 1), (1081, 1023)])
def test_depth_epilogue_v2_is_bit_identical(H, W):
    """DEPTHPRO_HBM_V2 (four output pixels per thread) must reproduce the default epilogue bit for bit, at aligned,
    ragged and tiny output sizes.  Runs the whole bf16 frame twice on the same input."""
    import depth_pro

    model = _experimental_model()
    g = torch.Generator(device=DEV).manual_seed(H * 10007 + W)
    x = torch.rand(3, H, W, device=DEV, generator=g) * 2 - 1
    try:
        _kb(8 | 0x2000, 64, 64, 0)
        a = model.infer(x)["depth"].clone()
        _kb(8 | 0x1000, 64, 64, 0)
        b = model.infer(x)["depth"].clone()
    finally:
        _kb(8 | 0x2000, 64, 64, 0)
    assert a.shape == b.shape
    assert torch.equal(a, b)


_MODEL = None


def _experimental_model():
    global _MODEL
    if _MODEL is None:
        import depth_pro

        _MODEL = depth_pro.DepthPro(device=torch.device(DEV), precision=torch.bfloat16).init_weights("stress", 1234)
    return _MODEL
