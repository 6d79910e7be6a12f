#!/usr/bin/env python3
"""Per-kernel SASS opcode histogram of libnnam_b200.so (cuobjdump -sass): the evidence that the contraction kernels are
tcgen05 / TMEM / TMA code (UTCHMMA = tcgen05.mma, LDTM = tcgen05.ld, UTMALDG / UTMASTG = TMA tensor load / store,
UBLKCP = cp.async.bulk, SYNCS = mbarrier) and not mma.sync (HMMA) dressed up.

  python scripts/sass_histogram.py > profiles/r02_sass_opcodes.txt
"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "nnacousticmodeling_b200", "libnnam_b200.so")
KEY = ["UTCHMMA", "UTCHMMA.2CTA", "LDTM", "UTMALDG", "UTMASTG", "UBLKCP", "UTCBAR", "SYNCS", "HMMA", "LDG", "STG", "LDS",
       "STS", "MUFU", "SHFL", "REDG", "ELECT"]


def main():
    out = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True, check=True).stdout
    filt = subprocess.run(["c++filt"], input=out, capture_output=True, text=True).stdout or out
    kernels, cur = collections.OrderedDict(), None
    for line in filt.splitlines():
        m = re.match(r"\s*Function : (.*)", line)
        if m:
            cur = kernels.setdefault(m.group(1).strip(), collections.Counter())
            continue
        m = re.match(r"\s*/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d+\s+)?([A-Z][A-Z0-9_.]*)", line)
        if m and cur is not None:
            op = m.group(1)
            cur[op.split(".")[0]] += 1
            if op.startswith("UTCHMMA") and ".2CTA" in op:
                cur["UTCHMMA.2CTA"] += 1
            cur["_total"] += 1
    total = collections.Counter()
    print(f"# {os.path.relpath(LIB, ROOT)}: {len(kernels)} kernel instances, cuobjdump -sass, nvcc "
          + subprocess.run(["nvcc", "--version"], capture_output=True, text=True).stdout.strip().splitlines()[-2])
    print("# columns: instructions, then the counts of the opcodes that identify the hardware path\n")
    groups = collections.OrderedDict()
    for name, c in kernels.items():
        base = re.sub(r"<.*", "", name.split("(")[0]).replace("nnam::", "").replace("void ", "")
        g = groups.setdefault(base, [0, collections.Counter()])
        g[0] += 1
        g[1].update(c)
        total.update(c)
    print(f"{'kernel (all template instances summed)':48s} {'inst':>5s} {'SASS':>8s} " + " ".join(f"{k:>8s}" for k in KEY))
    for base, (n, c) in groups.items():
        print(f"{base[:48]:48s} {n:5d} {c['_total']:8d} " + " ".join(f"{c[k]:8d}" for k in KEY))
    print(f"{'TOTAL':48s} {len(kernels):5d} {total['_total']:8d} " + " ".join(f"{total[k]:8d}" for k in KEY))
    if total["HMMA"]:
        print("\nWARNING: HMMA (mma.sync) instructions present", file=sys.stderr)


if __name__ == "__main__":
    main()
