"""What does the REFERENCE's own reduced-precision mode (model.half(), depth_pro.py:122-123) do to the result?  The oracle
(CPU restatement, pinned bit-exact to the reference) is run with every weight and activation in fp16 and in bf16 and
compared with its fp32 run on the bench frame / recipe-B weights.  Output: profiles/r2_reduced_precision_probe.json.
Context: the B200 engine maps precision=torch.half onto its bf16 mode (bf16 storage, fp32 accumulate / residual stream /
statistics), see INTEGRATION.md."""
import os, sys, time, json
ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..")
sys.path.insert(0, os.path.join(ROOT, "oracle")); sys.path.insert(0, os.path.join(ROOT, "ml-depth-pro-video_b200"))
import torch
import depthpro_oracle as O
from depth_pro import weights
torch.set_num_threads(os.cpu_count())
sd = weights.stress_init(1234)
x = O.synthetic_image_1536(1)
t0=time.time(); ref = O.infer(sd, x); t1=time.time()
print("fp32 s", t1-t0, flush=True)
res={}
for name, dt in (("fp16", torch.float16), ("bf16", torch.bfloat16)):
    sdh = {k: v.to(dt) for k, v in sd.items()}
    t0=time.time()
    try:
        out = O.infer(sdh, x.to(dt))
    except Exception as e:
        print(name, "failed", e, flush=True); continue
    d = out["depth"].float()
    ok = (ref["depth"] < 1e4-1) & (d < 1e4-1)
    rel = ((d-ref["depth"]).abs()/ref["depth"])[ok]
    q = torch.quantile(rel[::4], torch.tensor([0.5,0.99]))
    f_rel = abs(float(out["focallength_px"])-float(ref["focallength_px"]))/float(ref["focallength_px"])
    res[name] = {"seconds": time.time()-t0, "median_abs_rel": float(q[0]), "p99": float(q[1]), "max": float(rel.max()), "f_px_rel": f_rel, "clamp_mismatch": float((~ok).float().mean())}
    print(name, res[name], flush=True)
json.dump(res, open(os.path.join(ROOT, 'profiles', 'r2_reduced_precision_probe.json'), 'w'), indent=1)
