# LayerNorm-folded ViT GEMMs: core tests, model parity, A/B bench, isolated kernel rates.
mkdir -p gpurun_out
( timeout 600 python -m pytest tests/test_gpu_cores.py -x -q -m gpu -p no:cacheprovider -k "layernorm or residual or tma_store" ) > gpurun_out/t_cores.log 2>&1; echo "cores exit $?"; tail -15 gpurun_out/t_cores.log
( timeout 900 python -m pytest tests/test_gpu_model.py tests/test_gpu_video.py -x -q -m gpu -p no:cacheprovider -s ) > gpurun_out/t_model.log 2>&1; echo "model exit $?"; tail -25 gpurun_out/t_model.log
( timeout 300 python scripts/kernel_bench.py qkv fc1+gelu proj+res fc2+res "qkv LN-folded" "fc1+gelu LN-folded" "proj+res +LN out" "fc2+res +LN out" layernorm ) > gpurun_out/kernel_bench2.log 2>&1; echo "kb exit $?"; head -12 gpurun_out/kernel_bench2.log
( DEPTHPRO_LN_FUSE=0 timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline ) > gpurun_out/bench_lnfuse0.json 2> gpurun_out/bench_lnfuse0.err; echo "fuse0 exit $?"
( timeout 300 python bench.py --steps 20 --warmup 5 ) > gpurun_out/bench_lnfuse1.json 2> gpurun_out/bench_lnfuse1.err; echo "fuse1 exit $?"
python - <<'PY'
import json
for f in ("lnfuse0","lnfuse1"):
    try:
        d=json.loads(open(f"gpurun_out/bench_{f}.json").read().strip().splitlines()[-1])
        print(f, "value", round(d["value"],2), "e2e", round(d["e2e"]["value"],2), "p50", round(d["p50_ms_per_frame"],3), d["clocks"], "roof", round(d["roofline"]["frac"],3), "launches", d["gpu_launches"])
        print("   kernels:", json.dumps(d.get("kernels")))
        print("   cpu:", json.dumps(d.get("cpu_baseline")))
    except Exception as e: print(f, "ERR", e); print(open(f"gpurun_out/bench_{f}.err").read()[-1500:])
PY
