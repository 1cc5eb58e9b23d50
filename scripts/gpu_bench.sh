mkdir -p gpurun_out
rm -f gpurun_out/summary.txt
runall() { name=$1; shift; timeout $1 python -m pytest "${@:2}" -m gpu -q -rA --no-header -p no:cacheprovider > gpurun_out/$name.log 2>&1; echo "$name exit $?" >> gpurun_out/summary.txt; }
runall t_kernels 400 tests/test_gpu_kernels.py
runall t_conv 300 tests/test_gpu_cores.py::test_conv3x3
timeout 900 python bench.py > gpurun_out/bench_default.json 2> gpurun_out/bench_default.err; echo "bench_default exit $?" >> gpurun_out/summary.txt
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_reference.json 2> gpurun_out/bench_reference.err; echo "bench_ref exit $?" >> gpurun_out/summary.txt
timeout 600 python bench.py --batch 4 --steps 10 --no-cpu-baseline > gpurun_out/bench_b4.json 2> gpurun_out/bench_b4.err; echo "bench_b4 exit $?" >> gpurun_out/summary.txt
timeout 600 python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/plain.log 2>&1 && \
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -s 1119 -c 2700 --csv --log-file gpurun_out/launches.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/ncu.log 2>&1; echo "ncu exit $?" >> gpurun_out/summary.txt
cat gpurun_out/summary.txt; cat gpurun_out/bench_default.json
