# round 2, call 7: red-zone + parity run of the kernel cases, then the ncu evidence of the round
mkdir -p gpurun_out; rm -f gpurun_out/summary.txt
( timeout 300 python scripts/sanitize_cases.py kernels infer ) > gpurun_out/redzone_cases.log 2>&1; echo "redzone exit $?" >> gpurun_out/summary.txt
tail -4 gpurun_out/redzone_cases.log
bash scripts/gpu_ncu_r2.sh 2>&1 | tail -15
cat gpurun_out/summary.txt
