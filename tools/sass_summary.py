#!/usr/bin/env python
"""Per-kernel SASS opcode summary of libepn_b200.so (what proves the Blackwell-native paths: UTCHMMA = tcgen05.mma,
LDTM = tcgen05.ld, UTMALDG = tensor-map TMA load, UBLKCP = bulk async copy, LDGSTS = cp.async, FFMA2 = packed fp32
FMA, REDG/RED = fp32 atomics).

    python tools/sass_summary.py > profiles/r2_sass_summary.txt
"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SO = os.path.join(ROOT, "epn_pointcloud_b200", "libepn_b200.so")
WATCH = ["UTCHMMA", "LDTM", "UTMALDG", "UBLKCP", "LDGSTS", "SYNCS", "FFMA2", "FFMA", "FMUL2", "FADD2", "HMMA", "RED", "ATOMG",
         "LDS", "STS", "LDG", "STG", "LDL", "STL", "BAR", "REDUX", "SHFL"]


def main():
    out = subprocess.run(["cuobjdump", "-sass", SO], capture_output=True, text=True, check=True).stdout
    kernels = collections.OrderedDict()
    cur = None
    for line in out.splitlines():
        m = re.match(r"\s*Function : (\S+)", line)
        if m:
            cur = subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout.strip()
            cur = cur.replace("(anonymous namespace)::", "")
            cur = re.sub(r"\(.*", "", cur).replace("void ", "")
            kernels[cur] = collections.Counter()
            continue
        m = re.match(r"\s*/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_]+)", line)
        if m and cur is not None:
            kernels[cur][m.group(1)] += 1
            kernels[cur]["_total"] += 1
    print("%-72s %7s  %s" % ("kernel", "instrs", "watched opcodes (static counts)"))
    for name, c in kernels.items():
        w = ["%s %d" % (op, sum(v for k, v in c.items() if k == op or (op in ("RED", "BAR") and k.startswith(op)))) for op in WATCH]
        w = [x for x in w if not x.endswith(" 0")]
        print("%-72s %7d  %s" % (name[:72], c["_total"], ", ".join(w)))


if __name__ == "__main__":
    main()
