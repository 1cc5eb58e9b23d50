"""A/B of L2 persistence for the fp32 residual stream (DEPTHPRO_L2_PERSIST_MB) on B200, one
process: isolated proj / fc2 (dp_kernel_bench kind 13) and the full frame, off vs set-asides of 48 / 64 / 96 MB.

    DEPTHPRO_VERBOSE=1 python scripts/l2_persist_ab.py > gpurun_out/l2_persist_ab.json
"""
import ctypes
import json
import os
import sys

ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..")
sys.path.insert(0, os.path.join(ROOT, "ml-depth-pro-video_b200"))

import numpy as np
import torch

import depth_pro
from depth_pro import _capi, synthetic

dev = torch.device("cuda:0")
lib = _capi.load()
h = ctypes.c_void_p()
_capi.check(lib.dp_engine_create(0, _capi.PREC_BF16, 1, ctypes.byref(h)))
T = 37 * 577
SIZES = (0, 48, 64, 96)


def kb(kind, M, N, K, iters=30, mb=None):
    if mb is not None:
        kind |= 0x800 if mb == 0 else 0x400
        iters |= mb << 16
    ms = ctypes.c_float()
    _capi.check(lib.dp_kernel_bench(h, kind, M, N, K, iters, ctypes.byref(ms)))
    return ms.value * 1e3


res = {"isolated_us": {}}
for name, N, K in (("proj+res +LN out", 1024, 1024), ("fc2+res +LN out", 1024, 4096)):
    row = {}
    for mb in SIZES + SIZES:
        us = kb(13, T, N, K, mb=mb)
        row[str(mb)] = min(row.get(str(mb), 1e9), round(us, 2))
    res["isolated_us"][name] = row
    print(name, row, file=sys.stderr, flush=True)

model = depth_pro.DepthPro(device=dev, precision=torch.bfloat16).init_weights("stress", 1234)
x = synthetic.synthetic_image_1536(1).to(dev)
gold = np.load(os.path.join(ROOT, "tests", "golden", "reference_outputs.npz"))
gd = torch.from_numpy(gold["depth_1536"])


def frame(mb, steps=15):
    kb(13, 256, 1024, 1024, 1, mb=mb)  # flips the process-wide switch
    for _ in range(4):
        pred = model.infer(x)
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(steps):
        pred = model.infer(x)
    b.record()
    torch.cuda.synchronize()
    ms = a.elapsed_time(b) / steps
    rel = ((pred["depth"][::16, ::16].cpu() - gd).abs() / gd).flatten()
    return {"l2_persist_mb": mb, "ms_per_frame": round(ms, 3), "frames_per_s": round(1e3 / ms, 2),
            "depth_median_abs_rel": float(rel.median())}


res["frame"] = [frame(mb) for mb in (0, 96, 0, 64, 0, 48, 0, 96)]
for r in res["frame"]:
    print(r, file=sys.stderr, flush=True)
kb(13, 256, 1024, 1024, 1, mb=0)
print(json.dumps(res))
