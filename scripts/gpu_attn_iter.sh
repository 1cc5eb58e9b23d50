mkdir -p gpurun_out
( timeout 300 python -m pytest tests/test_gpu_cores.py -x -q -m gpu -p no:cacheprovider -k "attention" ) > gpurun_out/t_attn.log 2>&1; echo "attn test exit $?"; tail -3 gpurun_out/t_attn.log
timeout 120 scripts/ubench/attn_prof > gpurun_out/attn_prof.log 2>&1; echo "prof exit $?"; cat gpurun_out/attn_prof.log
( timeout 120 python scripts/kernel_bench.py "attention 37 seq" ) 2>&1 | head -2
