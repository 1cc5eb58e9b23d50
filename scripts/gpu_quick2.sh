mkdir -p gpurun_out
rm -f gpurun_out/summary.txt
runall() { name=$1; shift; timeout $1 python -m pytest "${@:2}" -m gpu -q -rA --no-header -p no:cacheprovider > gpurun_out/$name.log 2>&1; echo "$name exit $?" >> gpurun_out/summary.txt; }
runall t_model 1200 tests/test_gpu_model.py -s
runall t_video 600 tests/test_gpu_video.py
runall t_kernels 600 tests/test_gpu_kernels.py
timeout 900 python bench.py --no-cpu-baseline > gpurun_out/bench_default.json 2> gpurun_out/bench_default.err; echo "bench_default exit $?" >> gpurun_out/summary.txt
cat gpurun_out/summary.txt; grep -E "fp32|bf16|tap|PASS|FAIL|Error|error" gpurun_out/t_model.log | head -40; tail -15 gpurun_out/t_video.log; tail -4 gpurun_out/t_kernels.log; python -c "
import json; d=json.load(open('gpurun_out/bench_default.json')); print(d['value'], d['ms_per_step'], d['clocks']); print(json.dumps(d['kernels'],indent=1))"
