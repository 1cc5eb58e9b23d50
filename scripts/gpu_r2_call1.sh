# round 2, call 1: -m gpu suite, default bench, experimental attention variants (one process each), epilogue v2
mkdir -p gpurun_out; rm -f gpurun_out/summary.txt
run() { name=$1; t=$2; shift 2; ( timeout $t "$@" ) > gpurun_out/$name.log 2>&1; echo "$name exit $?" >> gpurun_out/summary.txt; }
run t_gpu 900 python -m pytest tests -q -m gpu -p no:cacheprovider
( timeout 300 python bench.py ) > gpurun_out/bench_default.json 2> gpurun_out/bench_default.err; echo "bench_default exit $?" >> gpurun_out/summary.txt
for v in 12 13 14; do
  ( ATTN_VARIANTS=$v ATTN_PINGPONG=1 timeout 120 python scripts/attn_variants.py --no-model ) > gpurun_out/attn_variant_$v.json 2> gpurun_out/attn_variant_$v.err; echo "attn_variant_$v exit $?" >> gpurun_out/summary.txt
done
run t_experimental 600 env DEPTHPRO_TEST_EXPERIMENTAL=1 python -m pytest tests/test_gpu_experimental.py -q -m gpu -p no:cacheprovider
( ATTN_VARIANTS=5,11 ATTN_PINGPONG=1 timeout 300 python scripts/attn_variants.py ) > gpurun_out/attn_variants.json 2> gpurun_out/attn_variants.err; echo "attn_variants exit $?" >> gpurun_out/summary.txt
( timeout 300 python scripts/kernel_bench.py ) > gpurun_out/kernel_bench.log 2>&1; echo "kernel_bench exit $?" >> gpurun_out/summary.txt
cat gpurun_out/summary.txt; tail -3 gpurun_out/t_gpu.log; tail -5 gpurun_out/t_experimental.log; tail -n 3 gpurun_out/attn_variant_1?.err | cut -c1-200
