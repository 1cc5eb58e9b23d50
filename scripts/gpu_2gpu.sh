mkdir -p gpurun_out
T="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511"
( timeout 400 $T bench.py --gpus 2 --steps 20 --warmup 5 ) > gpurun_out/bench_2gpu.json 2> gpurun_out/bench_2gpu.err; echo "2gpu exit $?"; head -c 900 gpurun_out/bench_2gpu.json; tail -3 gpurun_out/bench_2gpu.err
( timeout 400 $T bench.py --gpus 2 --workload clip1080p --steps 12 --warmup 3 ) > gpurun_out/bench_clip1080p_2gpu.json 2> gpurun_out/bench_clip1080p_2gpu.err; echo "clip 2gpu exit $?"; head -c 1200 gpurun_out/bench_clip1080p_2gpu.json; tail -3 gpurun_out/bench_clip1080p_2gpu.err
( timeout 300 $T bench.py --gpus 2 --impl reference --steps 1 --warmup 1 ) > gpurun_out/bench_ref_2gpu.json 2> gpurun_out/bench_ref_2gpu.err; echo "ref 2gpu exit $?"; head -c 600 gpurun_out/bench_ref_2gpu.json
