#!/usr/bin/env python
"""BASELINE configs[4]: synthetic sweep of ONE stride-1 InterSO3Conv forward,
N in {1k,4k,16k}, K in {32,64}, A in {12,60}, C in {32,128}:
points on the unit sphere surface, radius = 2*sqrt(K/N), sigma = 0.5*radius^2, feats randn, seed 100+i,
batch sized so that in+out bytes >= 512 MB (> 126 MB L2).  Prints one JSON line per shape with clouds/s and
the achieved fraction of the HBM and tensor rooflines of the FUSED layer accounting:
    bytes = 4*C*N*A (feats in) + 12*N (xyz) + 4*C*N*A (out) + 4*N*K (idx)
    flops = 2*C*N*A*KS*K + 2*C*C*KS*N*A + 11*N*A*KS*K
A=12 ("first 12 anchors") is not a reference configuration; every shape runs the fused kernel (rows of up to 128
slots; anchor subsets use the 60-anchor lane geometry with the other lanes dead, i.e. at 1/5 of the lane efficiency).

    python tools/sweep.py [--quick] > profiles/sweep.jsonl
"""
import argparse
import itertools
import json
import math
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import epn_pointcloud_b200 as E  # noqa: E402
from epn_pointcloud_b200 import functional as L  # noqa: E402

KS = 24


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--quick", action="store_true", help="N in {1k,4k} only")
    ap.add_argument("--iters", type=int, default=3)
    args = ap.parse_args()
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except OSError:
        pass
    hbm, tf = peaks.get("hbm_gbs", 6650.0), peaks.get("bf16_tflops", 1590.0)
    dev = "cuda:0"
    ns = (1024, 4096) if args.quick else (1024, 4096, 16384)
    for i, (n, k, a, c) in enumerate(itertools.product(ns, (32, 64), (12, 60), (32, 128))):
        radius = 2.0 * math.sqrt(k / n)
        sigma = 0.5 * radius * radius
        per_cloud = 2 * 4 * c * n * a
        b = max(1, min(64, -(-512 * 2 ** 20 // per_cloud)))
        torch.manual_seed(100 + i)
        conv = E.InterSO3Conv(c, c, 1, 1, radius, sigma, k, lazy_sample=True, kanchor=60).to(dev)
        if a != 60:
            conv.anchors = torch.from_numpy(L.get_anchors(60)[:a].copy()).to(dev)
        g = torch.Generator().manual_seed(100 + i)
        xyz = torch.randn(b, 3, n, generator=g)
        xyz = (xyz / xyz.norm(dim=1, keepdim=True)).to(dev)
        feats = torch.randn(b, c, n, a, generator=g).to(dev)
        x = E.SphericalPointCloud(xyz, feats, None)
        with torch.no_grad():
            for _ in range(2):
                conv(x)
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(args.iters):
                conv(x)
            e1.record()
            torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / args.iters
        nbytes = b * (4.0 * c * n * a * 2 + 12.0 * n + 4.0 * n * k)
        flops = b * (2.0 * c * n * a * KS * k + 2.0 * c * c * KS * n * a + 11.0 * n * a * KS * k)
        print(json.dumps({"N": n, "K": k, "A": a, "C": c, "batch": b, "ms": round(ms, 3),
                          "clouds_per_s": round(b / (ms * 1e-3), 1),
                          "fused_GBps": round(nbytes / ms / 1e6, 1), "frac_hbm": round(nbytes / ms / 1e6 / hbm, 4),
                          "TFLOPs": round(flops / ms / 1e9, 2), "frac_bf16_burst": round(flops / ms / 1e9 / tf, 4),
                          "path": "fused" if a == 60 else "fused (%d of 60 anchor lanes live)" % a}), flush=True)
        del conv, feats, xyz, x
        torch.cuda.empty_cache()


if __name__ == "__main__":
    main()
