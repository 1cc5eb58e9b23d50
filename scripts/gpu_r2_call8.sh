# round 2, call 8 (1 GPU): new tests (resize v2, two threads), attention phase profiles, kernel bench
mkdir -p gpurun_out; rm -f gpurun_out/summary.txt
run() { name=$1; t=$2; shift 2; ( timeout $t "$@" ) > gpurun_out/$name.log 2>&1; echo "$name exit $?" >> gpurun_out/summary.txt; }
run t_new 900 python -m pytest tests/test_gpu_model.py tests/test_gpu_kernels.py -q -m gpu -p no:cacheprovider -k "two_engines or resize or epilogue_v2 or interpolation"
for v in 13 12 5; do ( DEPTHPRO_ATTN_EXP=$v timeout 120 scripts/ubench/attn_prof ) > gpurun_out/attn_phase_v$v.log 2>&1; echo "attn_prof $v exit $?" >> gpurun_out/summary.txt; done
( timeout 300 python scripts/kernel_bench.py ) > gpurun_out/kernel_bench.log 2>&1; echo "kernel_bench exit $?" >> gpurun_out/summary.txt
cat gpurun_out/summary.txt; grep -E "passed|failed|error" gpurun_out/t_new.log | tail -3; grep -E "^FAILED|^E  " gpurun_out/t_new.log | cut -c1-300 | head; cat gpurun_out/attn_phase_v13.log; grep -E "resize|epilogue|attention" gpurun_out/kernel_bench.log | head
