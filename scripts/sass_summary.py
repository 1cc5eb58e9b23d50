"""SASS evidence for the tcgen05 / TMEM / TMA claims: instruction counts per kernel of the
in-tree library, from `cuobjdump -sass`.

    python scripts/sass_summary.py > profiles/sass_summary_r2.txt

UTCHMMA = tcgen05.mma (".2CTA" = cta_group::2), LDTM / STTM = tcgen05.ld / tcgen05.st, UTMALDG / UTMASTG = TMA
cp.async.bulk.tensor load / store, UTMAPF = tensor-map prefetch, UBLKPF = cp.async.bulk.prefetch.L2, UTCBAR = tcgen05.commit
(-> mbarrier), SYNCS = mbarrier ops, MUFU.EX2 = ex2.approx, HMMA = legacy mma.sync (debug attention only).
"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..")
LIB = os.path.join(ROOT, "ml-depth-pro-video_b200", "depth_pro", "libdepthpro_b200.so")
PATTERNS = ["UTCHMMA.2CTA", "UTCHMMA", "LDTM", "STTM", "UTMALDG", "UTMASTG", "UTMAPF", "UBLKPF", "UTCBAR", "SYNCS",
            "MUFU.EX2", "HMMA", "FFMA2", "F2FP"]


def main():
    sass = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True, check=True).stdout
    demangle = {}
    counts = collections.OrderedDict()
    arch = set(re.findall(r"arch = (sm_\w+)", sass))
    fn = None
    for line in sass.splitlines():
        m = re.match(r"\s*Function : (\S+)", line)
        if m:
            fn = m.group(1)
            counts[fn] = collections.Counter()
            continue
        if fn is None:
            continue
        m = re.match(r"\s*/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", line)
        if not m:
            continue
        op = m.group(1)
        counts[fn]["_total"] += 1
        for p in PATTERNS:
            if op.startswith(p):
                counts[fn][p] += 1
                if p == "UTCHMMA.2CTA":
                    continue
                break
    names = list(counts)
    dem = subprocess.run(["cu++filt"] + names, capture_output=True, text=True).stdout.splitlines() if names else []
    for n, d in zip(names, dem):
        d = re.sub(r"^void ", "", d).replace("dp::<unnamed>::", "").replace("(int)", "")
        demangle[n] = d[: d.index(">(") + 1] if ">(" in d else d.split("(")[0]
    print(f"# {os.path.relpath(LIB, ROOT)}: architectures {sorted(arch)}; {len(names)} kernels")
    print("# UTCHMMA = tcgen05.mma (.2CTA = cta_group::2), LDTM/STTM = tcgen05.ld/st, UTMALDG/UTMASTG = TMA tensor load/store,")
    print("# UBLKPF = bulk L2 prefetch, UTCBAR = tcgen05.commit, SYNCS = mbarrier, MUFU.EX2 = ex2.approx, HMMA = mma.sync (debug attention)")
    cols = [p for p in PATTERNS]
    print(f"{'kernel':72s} {'instr':>7s} " + " ".join(f"{c:>12s}" for c in cols))
    tot = collections.Counter()
    for n in names:
        c = counts[n]
        if not any(c[p] for p in PATTERNS[:9]) and c["MUFU.EX2"] == 0:
            continue  # plain SIMT helper kernels: listed in the total only
        # UTCHMMA counts every form; .2CTA is the cta_group::2 subset
        print(f"{demangle.get(n, n)[:72]:72s} {c['_total']:7d} " + " ".join(f"{c[p]:12d}" for p in cols))
        tot.update(c)
    print(f"{'TOTAL (kernels listed)':72s} {tot['_total']:7d} " + " ".join(f"{tot[p]:12d}" for p in cols))
    lib_tot = collections.Counter()
    for c in counts.values():
        lib_tot.update(c)
    print(f"{'TOTAL (whole library)':72s} {lib_tot['_total']:7d} " + " ".join(f"{lib_tot[p]:12d}" for p in cols))


if __name__ == "__main__":
    main()
