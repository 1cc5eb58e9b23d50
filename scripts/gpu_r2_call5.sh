# round 2, call 5: CUDA graph + side streams A/B (same process order alternated by separate runs), then the whole suite
mkdir -p gpurun_out; rm -f gpurun_out/summary.txt
run() { name=$1; t=$2; shift 2; ( timeout $t "$@" ) > gpurun_out/$name.log 2>&1; echo "$name exit $?" >> gpurun_out/summary.txt; }
run t_quick 600 python -m pytest tests/test_gpu_model.py -q -m gpu -p no:cacheprovider -x -k "bf16_vs_oracle or batch_is or layernorm_fold_matches or non_default or weight_edits or edge_shapes"
for rep in 1 2; do
for cfg in "1 1" "0 0" "0 1" "1 0"; do
  set -- $cfg
  ( DEPTHPRO_GRAPH=$1 DEPTHPRO_STREAMS=$2 timeout 300 python bench.py --no-video --no-cpu-baseline --steps 30 ) > gpurun_out/bench_g$1_s$2_r$rep.json 2> gpurun_out/bench_g$1_s$2_r$rep.err
  echo "graph=$1 streams=$2 rep=$rep: $(python -c "import json,sys; d=json.load(open('gpurun_out/bench_g$1_s$2_r$rep.json')); print(round(d['value'],2), 'fps  e2e', round(d['e2e']['value'],2), ' launches', d['gpu_launches'], ' clk', d['clocks']['sm_mhz'], ' gemm frac', round(d['roofline']['frac'],4))" 2>&1 | tail -1)" >> gpurun_out/summary.txt
done; done
run t_gpu 1500 python -m pytest tests -q -m gpu -p no:cacheprovider
cat gpurun_out/summary.txt; grep -E "passed|failed|error" gpurun_out/t_quick.log gpurun_out/t_gpu.log | tail -5; grep -E "^FAILED|^E  " gpurun_out/t_quick.log gpurun_out/t_gpu.log | cut -c1-300 | head -30; tail -3 gpurun_out/bench_g1_s1_r1.err
