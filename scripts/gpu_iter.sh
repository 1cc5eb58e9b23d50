mkdir -p gpurun_out
rm -f gpurun_out/summary.txt
timeout 300 python scripts/kernel_bench.py > gpurun_out/kernel_bench.log 2>&1; echo "kb exit $?" >> gpurun_out/summary.txt
runall() { name=$1; shift; timeout $1 python -m pytest "${@:2}" -m gpu -q -rA --no-header -p no:cacheprovider > gpurun_out/$name.log 2>&1; echo "$name exit $?" >> gpurun_out/summary.txt; }
runall t_cores 600 tests/test_gpu_cores.py
runall t_model 1200 tests/test_gpu_model.py -s
timeout 900 python bench.py --no-cpu-baseline > gpurun_out/bench_default.json 2> gpurun_out/bench_default.err; echo "bench_default exit $?" >> gpurun_out/summary.txt
cat gpurun_out/summary.txt; head -20 gpurun_out/kernel_bench.log; tail -3 gpurun_out/t_cores.log; grep -E "fp32|bf16" gpurun_out/t_model.log | head -10
python -c "
import json; d=json.load(open('gpurun_out/bench_default.json')); print(d['value'], d['ms_per_step'], d['clocks']); print(json.dumps(d['kernels']))"
