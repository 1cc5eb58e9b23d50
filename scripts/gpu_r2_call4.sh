# round 2, call 4: whole -m gpu suite after the API additions (bicubic, fov configs, colorize range, idempotent finalize)
mkdir -p gpurun_out; rm -f gpurun_out/summary.txt
run() { name=$1; t=$2; shift 2; ( timeout $t "$@" ) > gpurun_out/$name.log 2>&1; echo "$name exit $?" >> gpurun_out/summary.txt; }
run t_gpu 1500 python -m pytest tests -q -m gpu -p no:cacheprovider -s
( timeout 300 python scripts/kernel_bench.py ) > gpurun_out/kernel_bench.log 2>&1; echo "kernel_bench exit $?" >> gpurun_out/summary.txt
cat gpurun_out/summary.txt; grep -E "passed|failed|error" gpurun_out/t_gpu.log | tail -5; grep -E "^FAILED|^E  " gpurun_out/t_gpu.log | head -40; grep -E "bicubic|fov=|frames/s" gpurun_out/t_gpu.log | head; grep -E "epilogue|colorize|resize" gpurun_out/kernel_bench.log | head -8
