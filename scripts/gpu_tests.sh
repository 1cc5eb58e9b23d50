mkdir -p gpurun_out
nvidia-smi > gpurun_out/smi.txt 2>&1; nproc > gpurun_out/nproc.txt; lscpu | head -20 >> gpurun_out/nproc.txt
run() { name=$1; shift; timeout $1 python -m pytest "${@:2}" -m gpu -q -rA -x --no-header -p no:cacheprovider > gpurun_out/$name.log 2>&1; echo "$name exit $?" >> gpurun_out/summary.txt; }
runall() { name=$1; shift; timeout $1 python -m pytest "${@:2}" -m gpu -q -rA --no-header -p no:cacheprovider > gpurun_out/$name.log 2>&1; echo "$name exit $?" >> gpurun_out/summary.txt; }
rm -f gpurun_out/summary.txt
runall t_kernels 400 tests/test_gpu_kernels.py
runall t_gemm_fp32 300 tests/test_gpu_cores.py::test_gemm_fp32
runall t_gemm_tc 300 tests/test_gpu_cores.py::test_gemm_bf16_tcgen05
runall t_conv 300 tests/test_gpu_cores.py::test_conv3x3
runall t_attn 300 tests/test_gpu_cores.py::test_attention
runall t_model 1200 tests/test_gpu_model.py -s
cat gpurun_out/summary.txt
