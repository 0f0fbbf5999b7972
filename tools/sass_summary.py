"""Opcode histogram of the Blackwell-specific SASS in libadb200.so, per kernel (cuobjdump -sass): UTC*MMA = tcgen05.mma,
UTMALDG / UTMASTG / UTMAREDG / UTMAPF = TMA loads / stores / reduce-adds / prefetch, LDTM = tcgen05.ld, UTCBAR = tcgen05.commit,
HMMA = legacy mma.sync (the U-Net's small attention), SYNCS = mbarrier, MUFU.TANH. Writes profiles/r2_sass_summary.txt."""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
so = os.path.join(ROOT, "audiodiffuser_b200", "libadb200.so")
out = subprocess.run(["cuobjdump", "-sass", so], capture_output=True, text=True, check=True).stdout
pat = re.compile(r"\b(UTC[A-Z]*MMA(?:\.[A-Z0-9_]+)*|UTMALDG(?:\.[A-Z0-9_]+)*|UTMASTG(?:\.[A-Z0-9_]+)*|UTMAREDG(?:\.[A-Z0-9_.]+)*|UTMAPF(?:\.[A-Z0-9_]+)*|"
                 r"UTCBAR(?:\.[A-Z0-9_]+)*|LDTM(?:\.[A-Z0-9_]+)*|STTM(?:\.[A-Z0-9_]+)*|HMMA(?:\.[A-Z0-9_]+)*|UBLKCP(?:\.[A-Z0-9_]+)*|MUFU\.TANH(?:\.[A-Z0-9_]+)*|"
                 r"SYNCS(?:\.[A-Z0-9_]+)*|UTCATOMSWS(?:\.[A-Z0-9_.]+)*)\b")
kern, hist = None, collections.OrderedDict()
for line in out.splitlines():
    m = re.search(r"Function : (\S+)", line)
    if m:
        kern = m.group(1)
        hist[kern] = collections.Counter()
        continue
    if kern:
        for op in pat.findall(line):
            hist[kern][op] += 1
lines = [f"cuobjdump -sass {os.path.relpath(so, ROOT)}  (sm_100a; nvcc 12.9) — Blackwell-specific opcodes per kernel", ""]
total = collections.Counter()
for k, h in hist.items():
    if not h:
        continue
    total.update(h)
    name = subprocess.run(["c++filt", k], capture_output=True, text=True).stdout.strip().split("(")[0]
    lines.append(name)
    lines.append("    " + ", ".join(f"{op} x{n}" for op, n in sorted(h.items())))
lines += ["", "whole library:", "    " + ", ".join(f"{op} x{n}" for op, n in sorted(total.items()))]
path = os.path.join(ROOT, "profiles", "r2_sass_summary.txt")
open(path, "w").write("\n".join(lines) + "\n")
print(path, len(lines), "lines")
