# round 2: the round-end sequence on one GPU -- smoke(), whole -m gpu suite, default bench, reference arm (short)
mkdir -p gpurun_out; rm -f gpurun_out/summary.txt
run() { name=$1; t=$2; shift 2; ( timeout $t "$@" ) > gpurun_out/$name.log 2>&1; echo "$name exit $?" >> gpurun_out/summary.txt; }
run smoke 600 python -c "import __graft_entry__ as g; g.smoke()"
run t_gpu 1800 python -m pytest tests -q -m gpu -p no:cacheprovider
( timeout 900 python bench.py ) > gpurun_out/bench_default.json 2> gpurun_out/bench_default.err; echo "bench_default exit $?" >> gpurun_out/summary.txt
( timeout 900 python bench.py --batch 16 --steps 5 --warmup 3 --no-cpu-baseline ) > gpurun_out/bench_batch16.json 2> gpurun_out/bench_batch16.err; echo "batch16 exit $?" >> gpurun_out/summary.txt
cat gpurun_out/summary.txt; tail -2 gpurun_out/smoke.log; grep -E "passed|failed|error" gpurun_out/t_gpu.log | tail -3; grep -E "^FAILED|^E  " gpurun_out/t_gpu.log | cut -c1-300 | head -20
python - <<'PY'
import json
for f in ("gpurun_out/bench_default.json","gpurun_out/bench_batch16.json"):
    try:
        d=[json.loads(l) for l in open(f) if l.startswith("{")][-1]
    except Exception as e:
        print(f, "unreadable", e); continue
    print(f, "value", round(d["value"],2), "e2e", round(d["e2e"]["value"],2), "launches", d["gpu_launches"], "roofline", d["roofline"] and round(d["roofline"]["frac"],4), "clk", d["clocks"]["sm_mhz"], d["clocks"]["reasons"])
    if d.get("video"):
        for k,v in d["video"].items(): print("  ", k, round(v["value"],2), round(v["e2e"]["value"],2))
    if d.get("hbm_kernels"):
        for k,v in d["hbm_kernels"].items(): print("  ", k, v["us"], "us", v["frac_of_hbm_peak"])
    print("  cpu", d.get("cpu_baseline"))
PY
tail -3 gpurun_out/bench_default.err
