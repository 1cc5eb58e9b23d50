mkdir -p gpurun_out; rm -f gpurun_out/summary.txt
( timeout 60 python -m pytest tests/test_gpu_cores.py -q -m gpu -k "attention or residual_l2" -p no:cacheprovider ) > gpurun_out/t_new.log 2>&1; echo "t_new exit $?" >> gpurun_out/summary.txt
( timeout 80 python bench.py --steps 20 --warmup 5 ) > gpurun_out/bench_v6.json 2> gpurun_out/bench_v6.err; echo "bench exit $?" >> gpurun_out/summary.txt
( timeout 60 python -m pytest tests/test_gpu_model.py -q -m gpu -x -k "bf16_vs_oracle or batch_is_bit" -p no:cacheprovider ) > gpurun_out/t_model.log 2>&1; echo "t_model exit $?" >> gpurun_out/summary.txt
cat gpurun_out/summary.txt; tail -3 gpurun_out/t_new.log; tail -3 gpurun_out/t_model.log; head -c 400 gpurun_out/bench_v6.json; tail -2 gpurun_out/bench_v6.err
