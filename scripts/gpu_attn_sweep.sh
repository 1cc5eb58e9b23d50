mkdir -p gpurun_out
timeout 120 python -m pytest tests/test_gpu_cores.py -m gpu -q -x -k attention -p no:cacheprovider 2>&1 | tail -3
for h in 1 0; do echo "pingpong $h"; DEPTHPRO_ATTN_PINGPONG=$h timeout 120 python scripts/kernel_bench.py "attention 37 seq" 2>&1 | head -1; done
