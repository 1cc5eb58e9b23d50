mkdir -p gpurun_out
for h in 0 1; do echo "pingpong $h"; DEPTHPRO_ATTN_PINGPONG=$h timeout 120 python scripts/kernel_bench.py "attention 37 seq" 2>&1 | head -1; done
timeout 300 python -m pytest tests/test_gpu_cores.py -m gpu -q -k attention -p no:cacheprovider 2>&1 | tail -3
