mkdir -p gpurun_out; rm -f gpurun_out/summary.txt
for v in 20 21; do ( ATTN_VARIANTS=$v ATTN_PINGPONG=1,0 timeout 120 python scripts/attn_variants.py --no-model ) > gpurun_out/attn_variant_$v.json 2> gpurun_out/attn_variant_$v.err; echo "attn_variant_$v exit $?" >> gpurun_out/summary.txt; done
( ATTN_VARIANTS=13 ATTN_PINGPONG=1 timeout 120 python scripts/attn_variants.py --no-model ) > gpurun_out/attn_variant_13b.json 2> gpurun_out/attn_variant_13b.err
for v in 21 20; do ( DEPTHPRO_ATTN_EXP=$v timeout 120 scripts/ubench/attn_prof ) > gpurun_out/attn_phase_v$v.log 2>&1; done
cat gpurun_out/summary.txt; tail -n 3 gpurun_out/attn_variant_2?.err gpurun_out/attn_variant_13b.err | cut -c1-260; cat gpurun_out/attn_phase_v21.log gpurun_out/attn_phase_v20.log
