mkdir -p gpurun_out; rm -f gpurun_out/summary.txt
( timeout 900 python -m pytest tests/test_gpu_cores.py tests/test_gpu_model.py -q -m gpu -p no:cacheprovider -s -k "fp16" ) > gpurun_out/t_fp16.log 2>&1; echo "t_fp16 exit $?" >> gpurun_out/summary.txt
( timeout 300 python - <<'PY'
import sys, torch, time
sys.path.insert(0, "ml-depth-pro-video_b200")
import depth_pro
from depth_pro import synthetic
dev = torch.device("cuda:0")
x = synthetic.synthetic_image_1536(1).to(dev)
for prec in (torch.bfloat16, torch.float16):
    m = depth_pro.DepthPro(device=dev, precision=prec).init_weights("stress", 1234)
    for _ in range(5): m.infer(x)
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(30): m.infer(x)
    b.record(); torch.cuda.synchronize()
    print(prec, "frames/s", round(30e3 / a.elapsed_time(b), 2), flush=True)
    del m
PY
) > gpurun_out/fp16_speed.log 2>&1; echo "speed exit $?" >> gpurun_out/summary.txt
cat gpurun_out/summary.txt; grep -E "passed|failed|error|fp16:|tap " gpurun_out/t_fp16.log | tail -12; grep -E "^E  |^FAILED" gpurun_out/t_fp16.log | cut -c1-250 | head; cat gpurun_out/fp16_speed.log | tail -4
