mkdir -p gpurun_out
rm -f gpurun_out/*.ncu-rep
timeout 600 python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/plain.log 2>&1 && \
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -s 1125 -c 800 --csv --log-file gpurun_out/launches.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/ncu.log 2>&1; echo "ncu list exit $?"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:gemm_tc_kernel -s 448 -c 5 -o gpurun_out/prof_gemm_vit python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/ncu2.log 2>&1; echo "ncu vit exit $?"
timeout 900 ncu --set full --clock-control none -k regex:gemm_tc_kernel -s 545 -c 51 --csv --page raw --log-file gpurun_out/ncu_decoder_raw.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/ncu4.log 2>&1; echo "ncu decoder exit $?"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:attention_tc -s 72 -c 1 -o gpurun_out/prof_attn python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/ncu3.log 2>&1; echo "ncu attn exit $?"
ls -la gpurun_out | head -30; du -sh gpurun_out
