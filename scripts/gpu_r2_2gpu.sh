# round 2: the driver's N = 2 launch line (torchrun, one rank per GPU): headline + video keys + multi_gpu_bit_identical
mkdir -p gpurun_out
( timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus 2 --steps 20 --warmup 5 ) > gpurun_out/bench_2gpu.json 2> gpurun_out/bench_2gpu.err; echo "bench 2gpu exit $?"
tail -c 2500 gpurun_out/bench_2gpu.json; tail -5 gpurun_out/bench_2gpu.err
( timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29518 bench.py --impl reference --gpus 2 --steps 2 --warmup 1 ) > gpurun_out/bench_ref_2gpu.json 2> gpurun_out/bench_ref_2gpu.err; echo "ref 2gpu exit $?"
cut -c1-400 gpurun_out/bench_ref_2gpu.json
( timeout 600 python -m pytest tests/test_sharding.py -q -p no:cacheprovider ) 2>&1 | tail -3
