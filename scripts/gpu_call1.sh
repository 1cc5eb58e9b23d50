# Round-1 session-2 call 1: PDL A/B on the default bench, video workloads, kernel micro-bench, GPU tests.
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.max.sm,power.limit --format=csv > gpurun_out/smi.txt 2>&1
( DEPTHPRO_PDL=0 timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline ) > gpurun_out/bench_pdl0.json 2> gpurun_out/bench_pdl0.err; echo "pdl0 exit $?"
( DEPTHPRO_PDL=1 timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline ) > gpurun_out/bench_pdl1.json 2> gpurun_out/bench_pdl1.err; echo "pdl1 exit $?"
python - <<'PY'
import json
for f in ("pdl0","pdl1"):
    try:
        d=json.loads(open(f"gpurun_out/bench_{f}.json").read().strip().splitlines()[-1])
        print(f, "value", round(d["value"],2), "e2e", round(d["e2e"]["value"],2), "p50", round(d["p50_ms_per_frame"],3), d["clocks"], "roof", round(d["roofline"]["frac"],3))
        print("   hbm:", json.dumps(d.get("hbm_kernels")))
    except Exception as e: print(f, "ERR", e)
PY
( timeout 300 python bench.py --workload clip1080p --steps 12 --warmup 3 ) > gpurun_out/bench_clip1080p.json 2> gpurun_out/bench_clip1080p.err; echo "clip exit $?"; head -c 1500 gpurun_out/bench_clip1080p.json; tail -3 gpurun_out/bench_clip1080p.err
( timeout 300 python bench.py --workload stream4k --steps 8 --warmup 3 ) > gpurun_out/bench_stream4k.json 2> gpurun_out/bench_stream4k.err; echo "4k exit $?"; head -c 1500 gpurun_out/bench_stream4k.json; tail -3 gpurun_out/bench_stream4k.err
( timeout 300 python scripts/kernel_bench.py ) > gpurun_out/kernel_bench.log 2>&1; echo "kb exit $?"; cat gpurun_out/kernel_bench.log | head -30
( time timeout 900 python -m pytest tests/ -x -q -m gpu -p no:cacheprovider ) > gpurun_out/pytest.log 2>&1; echo "pytest exit $?"; tail -5 gpurun_out/pytest.log
