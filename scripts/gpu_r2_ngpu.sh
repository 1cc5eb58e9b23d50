# round 2: the driver's N-GPU launch line (torchrun, one rank per GPU): headline + video keys + multi_gpu_bit_identical
#   gpurun --gpus N --timeout 1200 -- 'N=8 bash scripts/gpu_r2_ngpu.sh'
N=${N:-2}
mkdir -p gpurun_out
( timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus $N --steps 20 --warmup 5 ) > gpurun_out/bench_${N}gpu.json 2> gpurun_out/bench_${N}gpu.err; echo "bench ${N}gpu exit $?"
python - <<PY
import json
d=[json.loads(l) for l in open("gpurun_out/bench_${N}gpu.json") if l.startswith("{")][-1]
print("N", d["n_gpus"], "value", round(d["value"],2), "e2e", round(d["e2e"]["value"],2), "bit-identical", d.get("multi_gpu_bit_identical"), "clk", d["clocks"])
for k,v in d["video"].items(): print("  ", k, round(v["value"],2), round(v["e2e"]["value"],2), v["frames_per_gpu"])
PY
tail -3 gpurun_out/bench_${N}gpu.err; nproc; nvidia-smi topo -m 2>/dev/null | head -12
