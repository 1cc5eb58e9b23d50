"""One-off, build container only: is the oracle PORT that `bench.py --impl reference` times on the GPU
box representative of the REFERENCE's own code?  Times full 1536^2 fp32 frames of (a) the unmodified
/root/reference/src/depth_pro model (`model.infer`, on the oracle/timm shim: timm itself is not in the image) and (b)
`oracle.infer` with the same random-init weights, same input, same torch thread count, alternating.

    python scripts/ref_vs_port_cpu.py [frames] > profiles/r2_reference_code_vs_port_cpu.json
"""
import json
import os
import sys
import time

ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..")
sys.path.insert(0, os.path.join(ROOT, "oracle"))
import torch

import depthpro_oracle as O
import reference_loader as RL

n = int(sys.argv[1]) if len(sys.argv) > 1 else 2
torch.set_num_threads(os.cpu_count())
model, _ = RL.build_reference_model()
sd = {k: v.detach().clone() for k, v in model.state_dict().items()}
x = O.synthetic_image_1536(1)
rows = []
with torch.no_grad():
    for i in range(n + 1):  # first pair = warm-up
        t0 = time.perf_counter()
        a = model.infer(x)
        t1 = time.perf_counter()
        b = O.infer(sd, x)
        t2 = time.perf_counter()
        rows.append({"reference_code_s": round(t1 - t0, 3), "oracle_port_s": round(t2 - t1, 3)})
        print(rows[-1], file=sys.stderr, flush=True)
same = bool(torch.equal(a["depth"], b["depth"]))
timed = rows[1:]
ref_s = sum(r["reference_code_s"] for r in timed) / len(timed)
port_s = sum(r["oracle_port_s"] for r in timed) / len(timed)
print(json.dumps({"what": "full 1536^2 fp32 frame on the build container's CPU, reference code (+ timm shim) vs oracle port",
                  "cores": os.cpu_count(), "torch_threads": torch.get_num_threads(), "frames_timed": len(timed),
                  "reference_code_s_per_frame": round(ref_s, 3), "oracle_port_s_per_frame": round(port_s, 3),
                  "port_over_reference": round(port_s / ref_s, 3), "outputs_bit_identical": same, "runs": rows}))
