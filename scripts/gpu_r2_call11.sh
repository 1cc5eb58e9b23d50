mkdir -p gpurun_out; rm -f gpurun_out/summary.txt
( timeout 600 python -m pytest tests/test_gpu_cores.py -q -m gpu -p no:cacheprovider -k attention ) > gpurun_out/t_attn.log 2>&1; echo "t_attn exit $?" >> gpurun_out/summary.txt
( ATTN_VARIANTS=13,22 ATTN_PINGPONG=1,0 FRAME_VARIANTS=13,22 timeout 400 python scripts/attn_variants.py ) > gpurun_out/attn_variants_22.json 2> gpurun_out/attn_variants_22.err; echo "attn_variants exit $?" >> gpurun_out/summary.txt
cat gpurun_out/summary.txt; tail -3 gpurun_out/t_attn.log; grep -E "^E  |^FAILED" gpurun_out/t_attn.log | head; cat gpurun_out/attn_variants_22.err | cut -c1-250
