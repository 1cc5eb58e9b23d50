"""A/B of the tcgen05 attention kernel's exp2 variants on B200, one process:

  1. parity of every variant (exp2 share on the FMA pipe 0 / 0 / 25 / 37.5 / 50 %, with and without the MUFU
     ping-pong) against fp64 SDPA on the bf16-rounded inputs, through dp_attention_test;
  2. isolated time of each over the frame's 37 sequences (dp_kernel_bench kind 4, random operands);
  3. the full frame (model.infer, bf16) with the default and with the fastest variant: frames/s from CUDA events
     and depth parity against the reference's recorded output (tests/golden/reference_outputs.npz).

    [ATTN_VARIANTS=0,1,5,6,7,8 ATTN_PINGPONG=1] python scripts/attn_variants.py [--no-model] > gpurun_out/attn_variants.json
"""
import ctypes
import json
import os
import sys

ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..")
sys.path.insert(0, os.path.join(ROOT, "ml-depth-pro-video_b200"))

import numpy as np
import torch
import torch.nn.functional as F

import depth_pro
from depth_pro import _capi, synthetic

dev = torch.device("cuda:0")
lib = _capi.load()
h = ctypes.c_void_p()
_capi.check(lib.dp_engine_create(0, _capi.PREC_BF16, 1, ctypes.byref(h)))
st = torch.cuda.current_stream(dev).cuda_stream
NAMES = {0: "scalar chain, all MUFU, strict ping-pong, P via smem (r1 v5)", 5: "fp32x2 chain, MUFU turn handed over at 12/16 (r1 default)",
         12: "5 + P through TMEM", 13: "12 + 25% poly"}
VARIANTS = [int(v) for v in os.environ.get("ATTN_VARIANTS", "0,5,12,13").split(",")]
PPS = [int(v) for v in os.environ.get("ATTN_PINGPONG", "1,0").split(",")]


def parity(v, pp, n):
    g = torch.Generator(device=dev).manual_seed(n)
    qkv = torch.randn(n, 577, 3072, device=dev, generator=g)
    out = torch.empty(n, 577, 1024, device=dev)
    backend = 1 | ((v + 1) << 8) | ((1 - pp) << 16)
    _capi.check(lib.dp_attention_test(h, backend, qkv.data_ptr(), out.data_ptr(), n, st))
    torch.cuda.synchronize()
    q, k, w = qkv.bfloat16().double().reshape(n, 577, 3, 16, 64).permute(2, 0, 3, 1, 4)
    ref = F.scaled_dot_product_attention(q, k, w).transpose(1, 2).reshape(n, 577, 1024)
    return float((out.double() - ref).abs().max() / ref.abs().max())


def timed(v, pp, iters=30):
    ms = ctypes.c_float()
    _capi.check(lib.dp_kernel_bench(h, 4, 37, 1 + v + 64 * (1 - pp), 0, iters, ctypes.byref(ms)))
    return ms.value * 1e3


res = {"variants": []}
for pp in PPS:
    for v in VARIANTS:
        e3 = parity(v, pp, 3)
        e37 = parity(v, pp, 37) if pp == 1 else None
        us = min(timed(v, pp), timed(v, pp))
        row = {"expv": v, "pingpong": pp, "name": NAMES[v], "relerr_n3": e3, "relerr_n37": e37, "us_37seq": round(us, 2),
               "TFLOPs": round(4.0 * 577 * 577 * 64 * 16 * 37 / us / 1e6, 1)}
        res["variants"].append(row)
        print(row, file=sys.stderr, flush=True)

ok = [r for r in res["variants"] if r["relerr_n3"] < 1.5e-2 and (r["relerr_n37"] is None or r["relerr_n37"] < 1.5e-2)]
best = min(ok, key=lambda r: r["us_37seq"]) if ok else None
res["best"] = best

if "--no-model" not in sys.argv and best is not None:
    model = depth_pro.DepthPro(device=dev, precision=torch.bfloat16).init_weights("stress", 1234)
    x = synthetic.synthetic_image_1536(1).to(dev)
    gold = np.load(os.path.join(ROOT, "tests", "golden", "reference_outputs.npz"))
    gd = torch.from_numpy(gold["depth_1536"])

    def frame(v, pp, steps=15):
        # set the process-wide variant through a 1-iteration kernel bench, then run whole frames
        timed(v, pp, 1)
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
        f_rel = abs(float(pred["focallength_px"]) - float(gold["f_px_1536"])) / float(gold["f_px_1536"])
        return {"expv": v, "pingpong": pp, "ms_per_frame": round(ms, 3), "frames_per_s": round(1e3 / ms, 2),
                "depth_median_abs_rel": float(rel.median()), "depth_max_abs_rel": float(rel.max()), "f_px_rel": f_rel}

    fv = [int(v) for v in os.environ.get("FRAME_VARIANTS", "5,13").split(",")]
    if best["expv"] not in fv:
        fv.append(best["expv"])
    res["frame"] = [frame(v, 1) for _ in range(3) for v in fv]   # alternating: the frame is power-capped and drifts
    for r in res["frame"]:
        print(r, file=sys.stderr, flush=True)
print(json.dumps(res))
