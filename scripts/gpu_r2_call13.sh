mkdir -p gpurun_out; rm -f gpurun_out/summary.txt
for v in 23 32; do ( ATTN_VARIANTS=$v ATTN_PINGPONG=1 timeout 120 python scripts/attn_variants.py --no-model ) > gpurun_out/attn_variant_$v.json 2> gpurun_out/attn_variant_$v.err; echo "attn_variant_$v exit $?" >> gpurun_out/summary.txt; done
( ATTN_VARIANTS=13,23 ATTN_PINGPONG=1 timeout 120 python scripts/attn_variants.py --no-model ) > gpurun_out/attn_variant_13d.json 2> gpurun_out/attn_variant_13d.err
cat gpurun_out/summary.txt; tail -n 4 gpurun_out/attn_variant_23.err gpurun_out/attn_variant_32.err gpurun_out/attn_variant_13d.err | cut -c1-300
