#!/usr/bin/env python
"""Per-kernel counts of the SASS mnemonics that prove which hardware paths libqck.so uses (TMA loads / stores,
FP64 tensor-core MMA, 256-bit stores, warp shuffles, cp.async, FP64 FMA).  Runs without a GPU:
  python tools/sass_summary.py > profiles/sass_summary.txt
"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "hardwareawareoptimalquantumcircuitcuttingandknitting_b200", "libqck.so")
PATTERNS = [
    ("UTMALDG", r"\bUTMALDG"), ("UTMASTG", r"\bUTMASTG"), ("DMMA", r"\bDMMA"), ("DFMA", r"\bDFMA"),
    ("STG.256", r"\bSTG\.[A-Z0-9.]*256"), ("LDG.256", r"\bLDG\.[A-Z0-9.]*256"), ("SHFL", r"\bSHFL"),
    ("LDGSTS", r"\bLDGSTS"), ("SYNCS", r"\bSYNCS"), ("BAR", r"\bBAR\."), ("LDS.128", r"\bLDS\.[A-Z0-9.]*128"),
    ("RED/ATOM", r"\b(RED|ATOMG|ATOMS)\b"),
]


def main():
    lib = sys.argv[1] if len(sys.argv) > 1 else LIB
    sass = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True, check=True).stdout
    arch = collections.Counter(re.findall(r"arch = (sm_\w+)", sass))
    counts, order, cur = collections.defaultdict(collections.Counter), [], None
    for line in sass.splitlines():
        m = re.match(r"\s*Function : (\S+)", line)
        if m:
            cur = subprocess.run(["c++filt", "-p", m.group(1)], capture_output=True, text=True).stdout.strip()
            order.append(cur)
            continue
        if cur is None or "/*" not in line:
            continue
        counts[cur]["instr"] += 1
        for name, pat in PATTERNS:
            if re.search(pat, line):
                counts[cur][name] += 1
    names = [n for n, _ in PATTERNS]
    print(f"# cuobjdump -sass {os.path.relpath(lib, ROOT)}: cubins {dict(arch)}, {len(order)} kernels")
    print(f"# {'kernel':<58}" + "".join(f"{n:>9}" for n in ["instr"] + names))
    total = collections.Counter()
    for k in sorted(order):
        c = counts[k]
        total.update(c)
        print(f"{k[:60]:<60}" + "".join(f"{c[n]:>9}" for n in ["instr"] + names))
    print(f"{'TOTAL':<60}" + "".join(f"{total[n]:>9}" for n in ["instr"] + names))
    print("# no tcgen05 / UTCMMA is expected: the 5th-generation tensor core has no FP64 kind; FP64 MMA is DMMA.")


if __name__ == "__main__":
    main()
