mkdir -p gpurun_out; rm -f gpurun_out/summary.txt
( timeout 900 python -m pytest tests/test_gpu_kernels.py tests/test_gpu_video.py tests/test_gpu_cores.py -q -m gpu -p no:cacheprovider -k "ground or fp16" ) > gpurun_out/t_ground.log 2>&1; echo "t_ground exit $?" >> gpurun_out/summary.txt
( timeout 300 python - <<'PY'
import sys, json, torch
sys.path.insert(0, ".")
sys.path.insert(0, "ml-depth-pro-video_b200")
import bench, depth_pro
from depth_pro import _capi
m = depth_pro.DepthPro(device=torch.device("cuda:0"), precision=torch.bfloat16).init_weights("stress", 1234)
m.infer(torch.zeros(3, 64, 64, device="cuda"))
print(json.dumps(bench.ground_kernels(_capi.load(), m, bench._peaks()[0]), indent=1))
PY
) > gpurun_out/ground_bench.log 2>&1; echo "ground bench exit $?" >> gpurun_out/summary.txt
cat gpurun_out/summary.txt; tail -4 gpurun_out/t_ground.log; grep -E "^E  |^FAILED" gpurun_out/t_ground.log | cut -c1-250 | head; tail -20 gpurun_out/ground_bench.log
