mkdir -p gpurun_out; rm -f gpurun_out/summary.txt
for v in 30 31; do ( ATTN_VARIANTS=$v ATTN_PINGPONG=1 timeout 120 python scripts/attn_variants.py --no-model ) > gpurun_out/attn_variant_$v.json 2> gpurun_out/attn_variant_$v.err; echo "attn_variant_$v exit $?" >> gpurun_out/summary.txt; done
( ATTN_VARIANTS=13 ATTN_PINGPONG=1 timeout 120 python scripts/attn_variants.py --no-model ) > gpurun_out/attn_variant_13c.json 2> gpurun_out/attn_variant_13c.err
cat gpurun_out/summary.txt; tail -n 4 gpurun_out/attn_variant_3?.err gpurun_out/attn_variant_13c.err | cut -c1-300
