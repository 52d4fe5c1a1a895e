#!/usr/bin/env python
"""Run ONE separable layer of the BASELINE classification backbone (forward + backward) so that ncu can
capture its kernels in isolation:

    ncu --set full --clock-control none --import-source on -k regex:'inter_group_tiles|umma_gemm|inter_scatter' \
        -c 6 -o gpurun_out/prof python tools/profile_layer.py --layer 3 --batch 8
"""
import argparse
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import epn_pointcloud_b200 as E  # noqa: E402
from epn_pointcloud_b200.blocks import cls_backbone_params  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--layer", type=int, default=3, help="0..6 = b0l0 .. b3l0")
    ap.add_argument("--batch", type=int, default=8)
    ap.add_argument("--iters", type=int, default=2)
    ap.add_argument("--forward-only", action="store_true", help="inference forward under no_grad (fused inter kernel only)")
    args = ap.parse_args()
    layers = [l["args"] for blk in cls_backbone_params(1024, 60) for l in blk]
    a = layers[args.layer]
    p_in = 1024
    for l in layers[:args.layer]:
        p_in = -(-p_in // l["stride"])
    torch.manual_seed(0)
    dev = "cuda:0"
    inter = E.InterSO3Conv(a["dim_in"], a["dim_out"], 1, a["stride"], a["radius"], a["sigma"], a["n_neighbor"],
                           lazy_sample=True, kanchor=60).to(dev)
    intra = E.IntraSO3Conv(a["dim_out"], a["dim_out"]).to(dev)
    g = torch.Generator().manual_seed(1)
    xyz = torch.randn(args.batch, 3, p_in, generator=g)
    xyz = (xyz / xyz.norm(dim=1, keepdim=True)).to(dev)
    feats = torch.randn(args.batch, a["dim_in"], p_in, 60, device=dev, requires_grad=True)
    for _ in range(args.iters):
        if args.forward_only:
            with torch.no_grad():
                inter(E.SphericalPointCloud(xyz, feats, None))
            continue
        _, _, _, y = inter(E.SphericalPointCloud(xyz, feats, None))
        z = intra(y)
        z.feats.square().mean().backward()
    torch.cuda.synchronize()
    print("layer %d: %d->%d, p_in %d, stride %d, K %d done" % (args.layer, a["dim_in"], a["dim_out"], p_in, a["stride"],
                                                             a["n_neighbor"]))


if __name__ == "__main__":
    main()
