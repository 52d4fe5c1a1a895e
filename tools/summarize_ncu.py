#!/usr/bin/env python
"""Summarise an `ncu --csv` launch list (metrics gpu__time_duration.sum [+ dram__bytes_read.sum,
dram__bytes_write.sum]) of ONE bench step into per-kernel and per-class totals.

    python tools/summarize_ncu.py gpurun_out/launches.csv [--json profiles/traffic_per_step.json]

The per-class DRAM bytes written with --json are what bench.py reports as `roofline.traffic`
(measured bytes per launch of the dominant kernel class, cold-cache/serialised ncu replay).
"""
import collections
import csv
import json
import re
import sys

CLASS_OF = [("umma_gemm", "channel_gemm"), ("umma_dw", "channel_gemm"), ("inter_bwd_fused", "inter_group_bwd_scatter"), ("inter_wt_steps", "split_convert"), ("inter_fused", "inter_fused_fwd"),
            ("intra_wt_tiles", "split_convert"), ("sgemm", "channel_gemm"), ("inter_group_tiles", "inter_group_fwd"), ("inter_group_direct", "inter_group_fwd"), ("inter_w_tiles", "split_convert"),
            ("inter_group_fwd", "inter_group_fwd"), ("inter_scatter", "inter_group_bwd_scatter"),
            ("inter_group_bwd", "inter_group_bwd_scatter"), ("intra_", "intra_group"), ("split_tiles", "split_convert"),
            ("norm_", "norm_act"), ("ball_query", "index_ops"), ("fps_kernel", "index_ops"), ("gather_", "index_ops")]


def to_base(value, unit):
    v = float(value.replace(",", ""))
    scale = {"ns": 1e-3, "us": 1.0, "usecond": 1.0, "ms": 1e3, "s": 1e6, "byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6,
             "Gbyte": 1e9}
    return v * scale.get(unit, 1.0)


def main():
    path = sys.argv[1]
    lines = [l for l in open(path) if not l.startswith("==")]
    per_kernel = collections.defaultdict(lambda: collections.defaultdict(float))
    counts = collections.Counter()
    for row in csv.DictReader(lines):
        name = re.sub(r"\(.*", "", row["Kernel Name"]).replace("void ", "")
        metric = row["Metric Name"]
        per_kernel[name][metric] += to_base(row["Metric Value"], row["Metric Unit"])
        if metric == "gpu__time_duration.sum":
            counts[name] += 1
    total_us = sum(v["gpu__time_duration.sum"] for v in per_kernel.values())
    print("%-58s %6s %11s %6s %10s %10s" % ("kernel", "n", "total us", "share", "rd MB", "wr MB"))
    per_class = collections.defaultdict(lambda: {"us": 0.0, "launches": 0, "dram_bytes": 0.0})
    for name, m in sorted(per_kernel.items(), key=lambda kv: -kv[1]["gpu__time_duration.sum"]):
        us = m["gpu__time_duration.sum"]
        rd, wr = m.get("dram__bytes_read.sum", 0.0), m.get("dram__bytes_write.sum", 0.0)
        print("%-58s %6d %11.1f %5.1f%% %10.1f %10.1f" % (name[:58], counts[name], us, 100 * us / total_us, rd / 1e6, wr / 1e6))
        for key, cls in CLASS_OF:
            if key in name and "epn::" in name:
                per_class[cls]["us"] += us
                per_class[cls]["launches"] += counts[name]
                per_class[cls]["dram_bytes"] += rd + wr
                break
    print("\nper class (this library only):")
    for cls, v in sorted(per_class.items(), key=lambda kv: -kv[1]["us"]):
        print("  %-26s %9.1f us  %5d launches  %9.1f MB DRAM" % (cls, v["us"], v["launches"], v["dram_bytes"] / 1e6))
    print("  all kernels of the step: %.1f us" % total_us)
    if "--json" in sys.argv:
        out = sys.argv[sys.argv.index("--json") + 1]
        json.dump({"source": path, "note": "DRAM bytes (read+write) per bench step and kernel class, ncu replay",
                   "classes": per_class}, open(out, "w"), indent=1)
        print("wrote", out)


if __name__ == "__main__":
    main()
