"""Condense the `ncu --set full --page raw --csv` exports of scripts/gpu_ncu_final.sh into the few columns the
design discussion uses, and write profiles/ncu_traffic.json (read by bench.py for `roofline.traffic`).

    python scripts/ncu_summarise.py gpurun_out/ncu_vit_block_raw.csv gpurun_out/ncu_decoder_raw.csv --tag r1_v5
"""
import argparse, csv, json, os, sys

KEEP = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "dram__throughput.avg.pct_of_peak_sustained_elapsed",
        "lts__t_sector_hit_rate.pct", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "launch__registers_per_thread", "launch__grid_size", "launch__block_size",
        "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum",
        "sm__cycles_elapsed.avg", "sm__cycles_elapsed.avg.per_second"]

UNIT = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "ns": 1e-3, "us": 1.0, "ms": 1e3, "s": 1e6}


def load(path):
    rows = [r for r in csv.reader(open(path, errors="replace")) if len(r) > 10]
    hdr = next(i for i, r in enumerate(rows) if "Kernel Name" in r)
    h, units, body = rows[hdr], rows[hdr + 1], rows[hdr + 2:]
    out = []
    for r in body:
        d = {"id": r[h.index("ID")], "kernel": r[h.index("Kernel Name")][:90]}
        for k in KEEP:
            if k in h:
                i = h.index(k)
                try:
                    v = float(r[i].replace(",", ""))
                except ValueError:
                    continue
                u = units[i]
                if u in UNIT:
                    v *= UNIT[u]            # bytes, microseconds
                d[k] = v
        out.append(d)
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("files", nargs="+")
    ap.add_argument("--tag", default="r1")
    ap.add_argument("--out-dir", default=os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "profiles"))
    a = ap.parse_args()
    allrows = []
    for f in a.files:
        rows = load(f)
        name = os.path.splitext(os.path.basename(f))[0]
        json.dump(rows, open(os.path.join(a.out_dir, f"{a.tag}_{name}_summary.json"), "w"), indent=1)
        allrows += rows
        for r in rows:
            t = r.get("gpu__time_duration.sum", 0)
            tr = r.get("dram__bytes_read.sum", 0) + r.get("dram__bytes_write.sum", 0)
            print(f"{r['id']:>5} {t:9.1f} us  dram {tr/1e6:8.1f} MB  tensor {r.get('sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed', 0):5.1f}%  {r['kernel'][:70]}")
    g = [r for r in allrows if "gemm_tc_kernel" in r["kernel"] and "dram__bytes_read.sum" in r]
    if g:
        tot = sum(r["dram__bytes_read.sum"] + r["dram__bytes_write.sum"] for r in g)
        json.dump({"gemm_tc_kernel": {"dram_bytes_per_launch_mean": tot / len(g), "launches_captured": len(g),
                                       "unit": "bytes", "source": f"ncu --set full, {a.tag}: " + ", ".join(os.path.basename(f) for f in a.files)}},
                  open(os.path.join(a.out_dir, "ncu_traffic.json"), "w"), indent=1)


if __name__ == "__main__":
    main()
