#!/usr/bin/env python
"""Per-kernel SASS evidence for the built library: counts of the mnemonics that prove (or disprove) a Blackwell-native kernel
(B200_PROFILING.md "What proves a Blackwell-native kernel").  Runs on the CPU box: cuobjdump -sass on
givepose_b200/lib/libgivepose_b200.so.  Writes profiles/sass_summary.md."""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SO = os.path.join(ROOT, "givepose_b200", "lib", "libgivepose_b200.so")
PATTERNS = [("UTCHMMA (tcgen05.mma)", r"\bUTC\w*MMA\b"), ("UTMALDG (TMA load)", r"\bUTMALDG"), ("UTMASTG/UBLKCP (TMA store / bulk)", r"\bUTMASTG|\bUBLKCP"),
            ("LDTM (tcgen05.ld)", r"\bLDTM"), ("SYNCS (mbarrier)", r"\bSYNCS"), ("HMMA (legacy mma.sync)", r"\bHMMA"),
            ("REDG.*F32x4 (vector fp32 reduction)", r"\bREDG\.E\.ADD\.F32x4|\bRED\.E\.ADD\.F32x4"), ("RED/ATOMG other", r"\bRED(G)?\.E\.(?!ADD\.F32x4)|\bATOMG"),
            ("ATOMS (shared atomics)", r"\bATOMS"), ("LDG.E.128", r"\bLDG\.E\.128"), ("LDS.128", r"\bLDS\.128"), ("SHFL", r"\bSHFL"),
            ("LDGSTS (cp.async)", r"\bLDGSTS")]


def main():
    txt = subprocess.run(["cuobjdump", "-sass", SO], capture_output=True, text=True, check=True).stdout
    counts, order, cur = collections.defaultdict(lambda: collections.Counter()), [], None
    for line in txt.splitlines():
        m = re.search(r"Function : (\S+)", line)
        if m:
            cur = subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout.strip()
            cur = re.sub(r"\(.*", "", cur).replace("void ", "").replace("gp::", "")
            order.append(cur)
            continue
        if cur is None or "/*" not in line:
            continue
        counts[cur]["instructions"] += 1
        for name, pat in PATTERNS:
            if re.search(pat, line):
                counts[cur][name] += 1
    # one row per kernel FAMILY (template instances summed, the instance count stated)
    fam = collections.OrderedDict()
    for k in order:
        f = re.sub(r"<.*", "", k)
        fam.setdefault(f, []).append(k)
    cols = [n for n, _ in PATTERNS]
    out = ["# SASS summary of `givepose_b200/lib/libgivepose_b200.so` (sm_100a)", "",
           "`python tools/sass_summary.py` (cuobjdump -sass, CUDA 12.9).  Counts are summed over the template instances of a kernel;",
           "PTX names never appear in SASS: `tcgen05.mma` = `UTC*MMA`, `tcgen05.ld` = `LDTM`, TMA = `UTMALDG` / `UTMASTG` / `UBLKCP`,",
           "`red.global.add.v4.f32` = `REDG.E.ADD.F32x4`, `mma.sync` would be `HMMA` (none).", "",
           "| kernel | instances | instructions | " + " | ".join(cols) + " |", "|---|---|---|" + "---|" * len(cols)]
    for f, ks in fam.items():
        tot = collections.Counter()
        for k in ks:
            tot.update(counts[k])
        out.append(f"| `{f}` | {len(ks)} | {tot['instructions']} | " + " | ".join(str(tot[c]) if tot[c] else "-" for c in cols) + " |")
    path = os.path.join(ROOT, "profiles", "sass_summary.md")
    open(path, "w").write("\n".join(out) + "\n")
    print("\n".join(out[:8] + out[8:][:60]))
    print("wrote", path)


if __name__ == "__main__":
    sys.exit(main())
