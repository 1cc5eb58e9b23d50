mkdir -p gpurun_out
timeout 300 python scripts/kernel_bench.py proj+res > gpurun_out/plain_kb.log 2>&1 && \
timeout 600 ncu --set full --clock-control none --import-source on -k regex:gemm_tc_kernel -s 3 -c 1 -o gpurun_out/prof_proj python scripts/kernel_bench.py proj+res > gpurun_out/ncu_kb.log 2>&1; echo "ncu exit $?"
tail -3 gpurun_out/ncu_kb.log
