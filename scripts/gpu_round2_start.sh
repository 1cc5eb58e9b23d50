# First GPU call of a new round: the whole -m gpu suite, the bench lines of every workload, the frame-level A/B of
# the attention variants that were only measured in isolation, then the ncu evidence.
#   gpurun --timeout 1500 -- 'bash scripts/gpu_round2_start.sh'
mkdir -p gpurun_out; rm -f gpurun_out/summary.txt
run() { name=$1; t=$2; shift 2; ( timeout $t "$@" ) > gpurun_out/$name.log 2>&1; echo "$name exit $?" >> gpurun_out/summary.txt; }
run t_gpu 900 python -m pytest tests -q -m gpu -p no:cacheprovider
( timeout 300 python bench.py ) > gpurun_out/bench_default.json 2> gpurun_out/bench_default.err; echo "bench_default exit $?" >> gpurun_out/summary.txt
( timeout 400 python bench.py --impl reference --steps 2 --warmup 1 ) > gpurun_out/bench_reference.json 2> gpurun_out/bench_reference.err; echo "bench_reference exit $?" >> gpurun_out/summary.txt
( timeout 300 python bench.py --workload clip1080p --no-cpu-baseline ) > gpurun_out/bench_clip1080p.json 2> gpurun_out/bench_clip1080p.err; echo "clip1080p exit $?" >> gpurun_out/summary.txt
( timeout 300 python bench.py --workload stream4k --no-cpu-baseline ) > gpurun_out/bench_stream4k.json 2> gpurun_out/bench_stream4k.err; echo "stream4k exit $?" >> gpurun_out/summary.txt
( timeout 300 python bench.py --batch 16 --steps 5 --warmup 3 --no-cpu-baseline ) > gpurun_out/bench_batch16.json 2> gpurun_out/bench_batch16.err; echo "batch16 exit $?" >> gpurun_out/summary.txt
# experimental variants one process each (a wrong barrier protocol traps after ~5 s and poisons the CUDA context)
for v in 12 13 14; do
  ( ATTN_VARIANTS=$v ATTN_PINGPONG=1 timeout 120 python scripts/attn_variants.py --no-model ) > gpurun_out/attn_variant_$v.json 2> gpurun_out/attn_variant_$v.err; echo "attn_variant_$v exit $?" >> gpurun_out/summary.txt
done
run t_experimental 600 env DEPTHPRO_TEST_EXPERIMENTAL=1 python -m pytest tests/test_gpu_experimental.py -q -m gpu -p no:cacheprovider
( ATTN_VARIANTS=5,11 ATTN_PINGPONG=1 timeout 300 python scripts/attn_variants.py ) > gpurun_out/attn_variants.json 2> gpurun_out/attn_variants.err; echo "attn_variants exit $?" >> gpurun_out/summary.txt
( timeout 300 python scripts/kernel_bench.py ) > gpurun_out/kernel_bench.log 2>&1; echo "kernel_bench exit $?" >> gpurun_out/summary.txt
cat gpurun_out/summary.txt; tail -3 gpurun_out/t_gpu.log; tail -8 gpurun_out/attn_variants.err | cut -c1-200; tail -n 2 gpurun_out/attn_variant_1?.err | cut -c1-200
bash scripts/gpu_ncu_final.sh
