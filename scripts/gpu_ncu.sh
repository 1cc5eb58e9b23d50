mkdir -p gpurun_out
timeout 600 python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/plain.log 2>&1 && \
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -s 1125 -c 1400 --csv --log-file gpurun_out/launches.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/ncu.log 2>&1; echo "ncu list exit $?"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:gemm_tc_kernel -s 20 -c 6 -o gpurun_out/prof_gemm python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/ncu2.log 2>&1; echo "ncu full exit $?"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:attention_tc -s 4 -c 1 -o gpurun_out/prof_attn python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/ncu3.log 2>&1; echo "ncu attn exit $?"
ls -la gpurun_out
