# HBM-kernel rewrites: kernel + video tests, micro-bench.
mkdir -p gpurun_out
( timeout 600 python -m pytest tests/test_gpu_kernels.py tests/test_gpu_video.py -x -q -m gpu -p no:cacheprovider ) > gpurun_out/t_kernels.log 2>&1; echo "kernels exit $?"; tail -15 gpurun_out/t_kernels.log
( timeout 600 python -m pytest tests/test_gpu_model.py -x -q -m gpu -p no:cacheprovider -k "1080p or errors" ) > gpurun_out/t_model2.log 2>&1; echo "model exit $?"; tail -5 gpurun_out/t_model2.log
( timeout 300 python scripts/kernel_bench.py ) > gpurun_out/kernel_bench3.log 2>&1; echo "kb exit $?"; head -28 gpurun_out/kernel_bench3.log
