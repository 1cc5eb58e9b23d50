# round 2, call 3: whole -m gpu suite, default bench (with video keys), reference arm quick check
mkdir -p gpurun_out; rm -f gpurun_out/summary.txt
run() { name=$1; t=$2; shift 2; ( timeout $t "$@" ) > gpurun_out/$name.log 2>&1; echo "$name exit $?" >> gpurun_out/summary.txt; }
run t_gpu 1200 python -m pytest tests -q -m gpu -p no:cacheprovider -s
( timeout 600 python bench.py ) > gpurun_out/bench_default.json 2> gpurun_out/bench_default.err; echo "bench_default exit $?" >> gpurun_out/summary.txt
( DEPTHPRO_REF_BUDGET_S=25 timeout 300 python bench.py --impl reference --steps 2 --warmup 1 ) > gpurun_out/bench_reference.json 2> gpurun_out/bench_reference.err; echo "bench_reference exit $?" >> gpurun_out/summary.txt
cat gpurun_out/summary.txt; grep -E "passed|failed|error" gpurun_out/t_gpu.log | tail -5; grep -E "^FAILED|^E  " gpurun_out/t_gpu.log | head -30; grep -E "frames/s|adversarial|seed 4321|outlier" gpurun_out/t_gpu.log | head; tail -c 1500 gpurun_out/bench_default.json; tail -3 gpurun_out/bench_default.err; cat gpurun_out/bench_reference.json | cut -c1-600
