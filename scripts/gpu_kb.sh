mkdir -p gpurun_out
timeout 300 python scripts/kernel_bench.py > gpurun_out/kernel_bench.log 2>&1; echo "kb exit $?"
cat gpurun_out/kernel_bench.log | head -20
timeout 900 python bench.py --no-cpu-baseline > gpurun_out/bench_default.json 2> gpurun_out/bench_default.err; echo "bench_default exit $?"
python -c "
import json; d=json.load(open('gpurun_out/bench_default.json')); print(d['value'], d['ms_per_step'], d['clocks']); print(json.dumps(d['kernels']))"
