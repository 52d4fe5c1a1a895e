#!/usr/bin/env python
"""Where a kernel's warps wait: the SASS instructions with the most stall samples of an `ncu --set full
--import-source on` report, with their dominant stall reasons, plus the sample share of the whole kernel per reason.

    python tools/ncu_stall_sites.py gpurun_out/x.ncu-rep [top_n] [launch] > profiles/rNN_stall_sites.txt

(launch = which kernel of a report holding several, default 0)
"""
import csv
import io
import subprocess
import sys


def main():
    path = sys.argv[1]
    top_n = int(sys.argv[2]) if len(sys.argv) > 2 else 30
    out = subprocess.run(["ncu", "-i", path, "--page", "source", "--csv", "--print-source", "sass"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    which = int(sys.argv[3]) if len(sys.argv) > 3 else 0
    hdrs = [i for i, r in enumerate(rows) if r and r[0] == "Address"]
    hdr_i = hdrs[which]
    end = hdrs[which + 1] - 1 if which + 1 < len(hdrs) else len(rows)
    print(rows[hdr_i - 1][:2] if hdr_i > 0 else "", "launch %d of %d" % (which, len(hdrs)))
    hdr, data = rows[hdr_i], [r for r in rows[hdr_i + 1:end] if len(r) == len(rows[hdr_i])]
    ci = {h: i for i, h in enumerate(hdr)}
    stalls = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
    total = sum(int(r[ci["# Samples"]]) for r in data)
    inst = sum(int(r[ci["Instructions Executed"]]) for r in data)
    print("samples %d, warp instructions %d" % (total, inst))
    per = {h: sum(int(r[ci[h]]) for r in data) for h in stalls}
    print("share of samples per reason:", ", ".join("%s %.1f%%" % (h[6:], 100.0 * v / max(total, 1)) for h, v in sorted(per.items(), key=lambda x: -x[1])[:8]))
    print("%6s %7s %10s  %-64s %s" % ("line", "samples", "executed", "instruction", "top reasons"))
    for r in sorted(data, key=lambda r: -int(r[ci["# Samples"]]))[:top_n]:
        st = sorted(((h[6:], int(r[ci[h]])) for h in stalls), key=lambda x: -x[1])[:2]
        print("%6d %7s %10s  %-64s %s" % (data.index(r), r[ci["# Samples"]], r[ci["Instructions Executed"]], r[ci["Source"]].strip()[:64], st))


if __name__ == "__main__":
    main()
