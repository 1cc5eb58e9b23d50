mkdir -p gpurun_out
for h in 16 14 12 10 8 6; do echo "handover $h"; DEPTHPRO_ATTN_HANDOVER=$h timeout 120 python scripts/kernel_bench.py "attention 37 seq" | head -1; done
timeout 300 python -m pytest tests/test_gpu_cores.py -m gpu -q -k attention -p no:cacheprovider 2>&1 | tail -2
