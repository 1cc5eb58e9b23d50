"""SM-partition experiment: the frame is power-capped, the attention kernel latency-bound.  Could a ViT
layer's GEMMs (on part of the SMs) and the attention of the other half of the batch (on the rest) share the chip?
  1. sustained fc1 / qkv throughput when the GEMM may use only L of the 148 SMs (is it power- or SM-bound?);
  2. one layer's GEMM chain + one attention launch, serial on the whole chip (kind 20) vs concurrent on two streams with
     the SMs split Lg / La (kind 21), each looped for ~2 s under the power cap.
    python scripts/sm_partition_probe.py > gpurun_out/sm_partition_probe.json
"""
import ctypes, json, os, sys
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "ml-depth-pro-video_b200"))
from depth_pro import _capi

lib = _capi.load()
h = ctypes.c_void_p()
_capi.check(lib.dp_engine_create(0, _capi.PREC_BF16, 1, ctypes.byref(h)))
T = 37 * 577


def bench(kind, M, N, K, iters, flags=0, hi=0):
    ms = ctypes.c_float()
    _capi.check(lib.dp_kernel_bench(h, kind | flags, M, N, K, (hi << 16) | iters, ctypes.byref(ms)))
    return ms.value * 1e3


def limits(gemm, attn):
    bench(5, 64, 0, 0, 1, 0x4000, gemm)   # sticky setters ride on a tiny LayerNorm bench
    bench(5, 64, 0, 0, 1, 0x8000, attn)


res = {"gemm_vs_sm_limit": [], "layer": []}
for L in (0, 132, 116, 100, 84, 68):
    limits(L, 0)
    row = {"sms": L or 148}
    for name, kind, N, K in (("fc1+gelu", 12, 4096, 1024), ("qkv", 11, 3072, 1024)):
        us = bench(kind, T, N, K, 50)
        us = bench(kind, T, N, K, max(200, int(1.5e6 / us)))
        row[name + "_us"] = round(us, 1)
        row[name + "_TF"] = round(2.0 * T * N * K / us / 1e6, 1)
    res["gemm_vs_sm_limit"].append(row)
    print(row, file=sys.stderr, flush=True)

limits(0, 0)
us = bench(20, T, 37, 0, 50)
serial = bench(20, T, 37, 0, max(200, int(2e6 / us)))
res["serial_layer_us"] = round(serial, 1)
print("serial layer (GEMM chain + attention, whole chip each):", round(serial, 1), "us", file=sys.stderr, flush=True)
for La in (24, 32, 40, 48, 56, 64):
    limits(148 - La, La)
    us = bench(21, T, 37, 0, 50)
    us = bench(21, T, 37, 0, max(200, int(2e6 / us)))
    row = {"gemm_sms": 148 - La, "attn_sms": La, "concurrent_layer_us": round(us, 1), "vs_serial": round(us / serial, 4)}
    res["layer"].append(row)
    print(row, file=sys.stderr, flush=True)
limits(0, 0)
print(json.dumps(res))
