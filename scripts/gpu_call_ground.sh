mkdir -p gpurun_out
( timeout 600 python -m pytest tests/test_gpu_kernels.py tests/test_gpu_video.py -x -q -m gpu -p no:cacheprovider -k "ground" ) > gpurun_out/t_ground.log 2>&1; echo "ground exit $?"; tail -30 gpurun_out/t_ground.log
