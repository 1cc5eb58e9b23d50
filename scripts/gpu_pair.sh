# pair-residual stream (EPI_RES16_LN): parity tests, isolated proj / fc2 timing (kind 13 fp32 stream vs kind 14 pair), frame A/B
mkdir -p gpurun_out
( timeout 900 python -m pytest tests/test_gpu_cores.py -q -k "producer or layernorm" -p no:cacheprovider ) > gpurun_out/t_pair.log 2>&1; echo "core tests exit $?"; tail -12 gpurun_out/t_pair.log | cut -c1-250
( timeout 1200 python -m pytest tests/test_gpu_model.py -q -x -s -k "pair_residual or bf16_vs_oracle or layernorm_fold or second_seed or zeros_init or fp16_mode" -p no:cacheprovider ) > gpurun_out/t_pair_model.log 2>&1; echo "model tests exit $?"; grep -E "pair vs|LN fold|passed|failed|Error|error" gpurun_out/t_pair_model.log | cut -c1-300 | tail -12
python - <<'PY' 2>&1 | tee gpurun_out/pair_bench.log
import ctypes, os, sys, time
sys.path.insert(0, "ml-depth-pro-video_b200")
import torch
from depth_pro import _capi
lib = _capi.load()
h = ctypes.c_void_p()
_capi.check(lib.dp_engine_create(0, _capi.PREC_BF16, 1, ctypes.byref(h)))
def t(kind, M, N, K, iters=40):
    ms = ctypes.c_float()
    _capi.check(lib.dp_kernel_bench(h, kind, M, N, K, iters, ctypes.byref(ms)))
    return ms.value * 1e3
for name, K in (("proj", 1024), ("fc2", 4096)):
    for rep in range(2):
        print(name, "fp32 stream (kind 13): %.1f us   pair (kind 14): %.1f us" % (t(13, 21349, 1024, K), t(14, 21349, 1024, K)), flush=True)
PY
for i in 1 2; do for p in 1 0; do echo "RES_PAIR=$p"; DEPTHPRO_RES_PAIR=$p timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-video 2>/dev/null | python -c "
import json,sys
d=[json.loads(l) for l in sys.stdin if l.startswith('{')][-1]
print('  value', round(d['value'],2), 'e2e', round(d['e2e']['value'],2), 'roofline', round(d['roofline']['frac'],4), 'clk', d['clocks']['sm_mhz'], 'kernels', {k:round(v['ms_per_step'],3) for k,v in d['kernels'].items()})
"; done; done 2>&1 | tee gpurun_out/pair_frame_ab.log
