"""A/B of the fp32-residual GEMMs' L2 prefetch (proj / fc2 epilogues) on B200, one process:
isolated proj / fc2 timings (dp_kernel_bench kinds 2 and 13, random operands) with the prefetch off / on, then the
full frame (model.infer, bf16) alternating off / on, with depth parity against the reference's recorded output.

    python scripts/res_prefetch_ab.py > gpurun_out/res_prefetch_ab.json
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
ON, OFF = 0x100, 0x200


def kb(kind, M, N, K, iters=30):
    ms = ctypes.c_float()
    _capi.check(lib.dp_kernel_bench(h, kind, M, N, K, iters, ctypes.byref(ms)))
    return ms.value * 1e3


res = {"isolated_us": {}}
for name, kind, N, K in (("proj+res (fp32 RMW)", 2, 1024, 1024), ("fc2+res (fp32 RMW)", 2, 1024, 4096),
                         ("proj+res +LN out", 13, 1024, 1024), ("fc2+res +LN out", 13, 1024, 4096)):
    row = {}
    for rep in range(2):
        for tag, bit in (("off", OFF), ("on", ON)):
            us = kb(kind | bit, T, N, K)
            row[tag] = min(row.get(tag, 1e9), round(us, 2))
    row["TFLOPs_off"] = round(2.0 * T * N * K / row["off"] / 1e6, 1)
    row["TFLOPs_on"] = round(2.0 * T * N * K / row["on"] / 1e6, 1)
    res["isolated_us"][name] = row
    print(name, row, file=sys.stderr, flush=True)

model = depth_pro.DepthPro(device=dev, precision=torch.bfloat16).init_weights("stress", 1234)
x = synthetic.synthetic_image_1536(1).to(dev)
gold = np.load(os.path.join(ROOT, "tests", "golden", "reference_outputs.npz"))
gd = torch.from_numpy(gold["depth_1536"])


def frame(bit, steps=15):
    kb(13 | bit, 1024, 1024, 1024, 1)  # flips the process-wide switch
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
    return {"prefetch": "on" if bit == ON else "off", "ms_per_frame": round(ms, 3), "frames_per_s": round(1e3 / ms, 2),
            "depth_median_abs_rel": float(rel.median()), "depth_max_abs_rel": float(rel.max())}


res["frame"] = [frame(b) for b in (OFF, ON, OFF, ON, OFF, ON)]
for r in res["frame"]:
    print(r, file=sys.stderr, flush=True)
print(json.dumps(res))
