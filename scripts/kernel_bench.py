"""Per-shape timings of the bf16 cores on B200 (CUDA events, scratch buffers)."""
import ctypes, sys, os, json
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "ml-depth-pro-video_b200"))
from depth_pro import _capi
lib = _capi.load()
h = ctypes.c_void_p()
_capi.check(lib.dp_engine_create(0, 1, 1, ctypes.byref(h)))
T = 37 * 577
cases = [("qkv", 0, T, 3072, 1024), ("fc1+gelu", 1, T, 4096, 1024), ("proj+res", 2, T, 1024, 1024),
         ("fc2+res", 2, T, 1024, 4096), ("patch_embed~", 0, 36 * 576, 1024, 768),
         ("conv768 256->256", 3, 768, 256, 256), ("conv384 256->256", 3, 384, 256, 256),
         ("conv768 256->128", 3, 768, 128, 256), ("conv96 1024->256", 3, 96, 256, 1024),
         ("attention 37 seq", 4, 37, 0, 0), ("layernorm", 5, T, 0, 0),
         ("fc1 no-gelu", 0, T, 4096, 1024),
         ("qkv LN-folded", 11, T, 3072, 1024), ("fc1+gelu LN-folded", 12, T, 4096, 1024),
         ("proj+res +LN out", 13, T, 1024, 1024), ("fc2+res +LN out", 13, T, 1024, 4096),
         # HBM-bound kernels either side of the network: (kind, H, W); bytes = algorithmic read + write
         ("resize 1080p u8->1536^2 f32", 6, 1080, 1920, 0), ("resize 4K u8->1536^2 f32", 6, 2160, 3840, 0),
         ("split+im2col 1536^2 -> 36x576x768 bf16", 7, 0, 0, 0),
         ("depth epilogue -> 1080p", 8, 1080, 1920, 0), ("depth epilogue -> 4K", 8, 2160, 3840, 0),
         ("unproject 4K (+rgb)", 9, 2160, 3840, 0), ("unproject 1080p (+rgb)", 9, 1080, 1920, 0),
         ("colorize 1080p", 10, 1080, 1920, 0)]


def hbm_bytes(kind, H, W):
    """Algorithmic bytes of the HBM-bound kernels."""
    img = 3 * 1536 * 1536
    return {6: H * W * 3 + img * 4, 7: img * 4 + 36 * 576 * 768 * 2, 8: 1536 * 1536 * 4 + H * W * 4,
            9: H * W * (4 + 3) + H * W * 24, 10: H * W * (4 + 3)}[kind]
out = {}
if len(sys.argv) > 1:
    cases = [c for c in cases if c[0] in sys.argv[1:]]
for name, kind, M, N, K in cases:
    ms = ctypes.c_float()
    _capi.check(lib.dp_kernel_bench(h, kind, M, N, K, 20, ctypes.byref(ms)))
    if kind <= 2 or 11 <= kind <= 13: fl = 2.0 * M * N * K
    elif kind == 3: fl = 2.0 * M * M * N * 9 * K
    elif kind == 4: fl = 4.0 * 577 * 577 * 64 * 16 * M
    else: fl = 0
    tf = fl / (ms.value * 1e-3) / 1e12 if fl else 0
    gbs = (M * 1024 * 6) / (ms.value * 1e-3) / 1e9 if kind == 5 else 0
    if 6 <= kind <= 10:
        gbs = hbm_bytes(kind, M, N) / (ms.value * 1e-3) / 1e9
    out[name] = {"us": round(ms.value * 1e3, 1), "TFLOP/s": round(tf, 1), "GB/s": round(gbs, 1)}
    print(f"{name:22s} {ms.value*1e3:9.1f} us  {tf:8.1f} TF/s {gbs:8.1f} GB/s", flush=True)
print(json.dumps(out))
