"""Per-kernel shares of one frame from an `ncu --metrics gpu__time_duration.sum` launch list (scripts/gpu_ncu_r2.sh).

    python scripts/launch_list_summary.py gpurun_out/launches.csv profiles/r2_launch_list_summary.json
"""
import csv
import json
import re
import sys

src, dst = sys.argv[1], sys.argv[2]
rows = [r for r in csv.reader(open(src, errors="replace")) if len(r) > 5]
hdr = next(i for i, r in enumerate(rows) if "Kernel Name" in r)
h = rows[hdr]
ik, iv, iu = h.index("Kernel Name"), h.index("Metric Value"), h.index("Metric Unit")
body = rows[hdr + 1:]
scale = {"ns": 1e-3, "us": 1.0, "ms": 1e3, "s": 1e6}
launches = [(r[ik], float(r[iv].replace(",", "")) * scale.get(r[iu], 1.0)) for r in body]
starts = [i for i, (k, _) in enumerate(launches) if "split_im2col" in k]
frame = launches[starts[0]:starts[1]]


def short(k):
    m = re.search(r"(\w+_kernel\w*)(<[^>]*>)?", k)
    return (m.group(1) + (m.group(2) or "")) if m else k[:40]


tot = sum(t for _, t in frame)
kern = {}
for k, t in frame:
    e = kern.setdefault(short(k), {"launches": 0, "us": 0.0})
    e["launches"] += 1
    e["us"] += t
for e in kern.values():
    e["us"] = round(e["us"], 1)
    e["share"] = round(e["us"] / tot, 4)
cls = {}
for k, e in kern.items():
    c = "gemm_tc_kernel" if k.startswith("gemm_tc_kernel") else ("attention_tc_kernel" if k.startswith("attention_tc") else "other")
    cls[c] = round(cls.get(c, 0.0) + e["us"] / tot, 4)
json.dump({"source": src + " (ncu --metrics gpu__time_duration.sum --clock-control none; DEPTHPRO_GRAPH=0 DEPTHPRO_STREAMS=0; "
                     "cold-cache, serialised launches: shares, not absolute times)",
           "launches_per_frame": len(frame), "sum_us_per_frame": round(tot, 1),
           "kernels": dict(sorted(kern.items(), key=lambda kv: -kv[1]["us"])), "class_shares": cls}, open(dst, "w"), indent=1)
print(len(frame), "launches per frame,", round(tot, 1), "us;", cls)
