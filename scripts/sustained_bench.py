"""Sustained (power-capped) rates of the ViT GEMM shapes: this engine's tcgen05 kernels vs cuBLAS (torch.matmul),
each looped for ~2 s, with the SM clock / power sampled from nvidia-smi meanwhile."""
import ctypes, json, os, subprocess, sys, tempfile, time
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "ml-depth-pro-video_b200"))
import torch
from depth_pro import _capi

lib = _capi.load()
h = ctypes.c_void_p()
_capi.check(lib.dp_engine_create(0, 1, 1, ctypes.byref(h)))
T = 37 * 577


class Smi:
    def __enter__(self):
        self.f = tempfile.NamedTemporaryFile("w+", suffix=".csv", delete=False)
        self.p = subprocess.Popen(["nvidia-smi", "-i", "0", "--query-gpu=clocks.sm,power.draw", "--format=csv,noheader,nounits",
                                   "-lms", "100"], stdout=self.f, stderr=subprocess.DEVNULL)
        return self

    def __exit__(self, *a):
        time.sleep(0.1)
        self.p.terminate(); self.p.wait(); self.f.flush(); self.f.seek(0)
        rows = [tuple(float(t) for t in l.split(",")) for l in self.f.read().splitlines() if l.count(",") == 1]
        rows = rows[len(rows) // 3:] or rows   # steady state
        self.mhz = sorted(r[0] for r in rows)[len(rows) // 2] if rows else None
        self.watts = sorted(r[1] for r in rows)[len(rows) // 2] if rows else None


def ours(kind, M, N, K, target_s=2.0):
    ms = ctypes.c_float()
    _capi.check(lib.dp_kernel_bench(h, kind, M, N, K, 50, ctypes.byref(ms)))
    iters = max(100, int(target_s / (ms.value * 1e-3)))
    with Smi() as s:
        _capi.check(lib.dp_kernel_bench(h, kind, M, N, K, iters, ctypes.byref(ms)))
    return ms.value, s.mhz, s.watts


def cublas(M, N, K, target_s=2.0):
    a = torch.randn(M, K, device="cuda", dtype=torch.bfloat16)
    w = torch.randn(N, K, device="cuda", dtype=torch.bfloat16)
    out = torch.empty(M, N, device="cuda", dtype=torch.bfloat16)
    for _ in range(20):
        torch.matmul(a, w.t(), out=out)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(50):
        torch.matmul(a, w.t(), out=out)
    e1.record(); torch.cuda.synchronize()
    iters = max(100, int(target_s / (e0.elapsed_time(e1) / 50 * 1e-3)))
    with Smi() as s:
        e0.record()
        for _ in range(iters):
            torch.matmul(a, w.t(), out=out)
        e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters, s.mhz, s.watts


cases = [("qkv", 11, T, 3072, 1024), ("fc1+gelu", 12, T, 4096, 1024), ("proj+res", 13, T, 1024, 1024),
         ("fc2+res", 13, T, 1024, 4096), ("conv768 256->256", 3, 768, 256, 256), ("attention", 4, 37, 0, 0)]
res = {}
for name, kind, M, N, K in cases:
    ms, mhz, w = ours(kind, M, N, K)
    fl = 2.0 * M * N * K if kind != 3 else 2.0 * M * M * N * 9 * K
    if kind == 4:
        fl = 4.0 * 577 * 577 * 64 * 16 * M
    r = {"ours_us": round(ms * 1e3, 1), "ours_TF": round(fl / ms / 1e9, 1), "ours_mhz": mhz, "ours_W": w,
         "ours_mJ": round(ms * w, 2) if w else None}
    if kind in (11, 12, 13):
        cms, cmhz, cw = cublas(M, N, K)
        r.update({"cublas_us": round(cms * 1e3, 1), "cublas_TF": round(fl / cms / 1e9, 1), "cublas_mhz": cmhz, "cublas_W": cw,
                  "cublas_mJ": round(cms * cw, 2) if cw else None})
    res[name] = r
    print(name, json.dumps(r), flush=True)
print(json.dumps(res))
