"""Build libdepthpro_b200.so (sm_100a) in-tree with nvcc.

    python ml-depth-pro-video_b200/build.py [--force]

Every csrc/*.cu is compiled to build/<name>.o (skipped when the source + headers hash is
unchanged) and linked into depth_pro/libdepthpro_b200.so.  nvcc cross-compiles without a GPU.
"""

from __future__ import annotations

import hashlib
import os
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
# two flavours from the same sources (csrc/common.cuh): 16-bit storage = bfloat16 (default) or IEEE half (-DDP_ACT_FP16)
FLAVOURS = {"bf16": ("build", "libdepthpro_b200.so", []), "fp16": ("build_fp16", "libdepthpro_b200_fp16.so", ["-DDP_ACT_FP16"])}
BUILD = os.path.join(HERE, "build")
OUT = os.path.join(HERE, "depth_pro", "libdepthpro_b200.so")
NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-std=c++17",
         "-Xcompiler", "-fPIC"]


def _digest(paths, extra=()):
    h = hashlib.sha256(" ".join(list(FLAGS) + list(extra)).encode())
    for p in sorted(paths):
        h.update(p.encode())
        with open(p, "rb") as f:
            h.update(f.read())
    return h.hexdigest()


def build(force: bool = False, verbose: bool = True) -> str:
    """Build every flavour; returns the path of the default (bf16) library."""
    for name in FLAVOURS:
        _build_flavour(name, force, verbose)
    return OUT


def _build_flavour(flavour: str, force: bool, verbose: bool) -> str:
    build_dir, lib_name, defs = FLAVOURS[flavour]
    BUILD = os.path.join(HERE, build_dir)
    OUT = os.path.join(HERE, "depth_pro", lib_name)
    os.makedirs(BUILD, exist_ok=True)
    headers = [os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith((".cuh", ".h"))]
    headers.append(os.path.join(os.path.dirname(HERE), "include", "depthpro_b200.h"))
    sources = sorted(os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith(".cu"))
    jobs = []
    objs = []
    for src in sources:
        obj = os.path.join(BUILD, os.path.basename(src)[:-3] + ".o")
        stamp = obj + ".sha"
        dig = _digest([src] + headers, defs)
        objs.append(obj)
        if not force and os.path.exists(obj) and os.path.exists(stamp) and open(stamp).read() == dig:
            continue
        jobs.append((src, obj, stamp, dig))

    def compile_one(job):
        src, obj, stamp, dig = job
        cmd = [NVCC] + FLAGS + defs + ["-c", src, "-o", obj]
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError(f"nvcc failed for {src}:\n{r.stdout}\n{r.stderr}")
        with open(stamp, "w") as f:
            f.write(dig)
        return src

    if jobs:
        with ThreadPoolExecutor(max_workers=min(8, len(jobs))) as ex:
            for done in ex.map(compile_one, jobs):
                if verbose:
                    print(f"[build] {flavour}: compiled {os.path.basename(done)}", flush=True)
    if jobs or force or not os.path.exists(OUT):
        cmd = [NVCC, "-shared", "-o", OUT] + objs + ["-gencode", "arch=compute_100a,code=sm_100a"]
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError(f"link failed:\n{r.stdout}\n{r.stderr}")
        # an unresolved symbol would only show up as a dlopen failure on the GPU box: check here
        chk = subprocess.run([sys.executable, "-c", f"import ctypes; ctypes.CDLL({OUT!r})"], capture_output=True, text=True)
        if chk.returncode != 0:
            os.remove(OUT)
            raise RuntimeError(f"{OUT} does not load:\n{chk.stderr[-2000:]}")
        if verbose:
            print(f"[build] linked {OUT}", flush=True)
    return OUT


if __name__ == "__main__":
    build(force="--force" in sys.argv)
