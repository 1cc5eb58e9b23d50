# round 2, call 2: new attention variants (parity + timing + frame A/B), hardened bf16 parity tests, MUFU f16x2 ubench
mkdir -p gpurun_out; rm -f gpurun_out/summary.txt
run() { name=$1; t=$2; shift 2; ( timeout $t "$@" ) > gpurun_out/$name.log 2>&1; echo "$name exit $?" >> gpurun_out/summary.txt; }
run t_attn 600 python -m pytest tests/test_gpu_cores.py -q -m gpu -p no:cacheprovider -k "attention" -x
for v in 15 16 17 18; do
  ( ATTN_VARIANTS=$v ATTN_PINGPONG=1 timeout 120 python scripts/attn_variants.py --no-model ) > gpurun_out/attn_variant_$v.json 2> gpurun_out/attn_variant_$v.err; echo "attn_variant_$v exit $?" >> gpurun_out/summary.txt
done
( ATTN_VARIANTS=5,12,13,15,16,17,18 ATTN_PINGPONG=1,0 FRAME_VARIANTS=5,13,16 timeout 400 python scripts/attn_variants.py ) > gpurun_out/attn_variants.json 2> gpurun_out/attn_variants.err; echo "attn_variants exit $?" >> gpurun_out/summary.txt
run t_model 900 python -m pytest tests/test_gpu_model.py -q -m gpu -p no:cacheprovider -s
( cd scripts/ubench && timeout 60 ./xu2 ) > gpurun_out/ubench_xu2.log 2>&1; echo "xu2 exit $?" >> gpurun_out/summary.txt
cat gpurun_out/summary.txt; tail -5 gpurun_out/t_attn.log; tail -n 2 gpurun_out/attn_variant_1?.err | cut -c1-250; tail -12 gpurun_out/attn_variants.err | cut -c1-250; grep -E "passed|failed|bf16|adversarial|outlier|border" gpurun_out/t_model.log | tail -20; cat gpurun_out/ubench_xu2.log
