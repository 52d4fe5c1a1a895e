#!/usr/bin/env python
"""Time the fused InterSO3Conv forward kernel on the six feature layers of the classification backbone (B = 32) under
the kernel's tuning knobs (EPN_FU_GATHER, EPN_FU_SPS), each combination in a fresh subprocess (the knobs are read
once per process).  CUDA events around 10 inference forwards of the layer, after 3 warm-ups.

    python tools/fused_sweep.py                    # all combinations -> gpurun_out/fused_sweep.json
    python tools/fused_sweep.py --one              # (internal) one configuration, prints a JSON line
"""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def one():
    import torch
    import epn_pointcloud_b200 as E
    from epn_pointcloud_b200.blocks import cls_backbone_params
    layers = [l["args"] for blk in cls_backbone_params(1024, 60) for l in blk]
    dev = "cuda:0"
    res = {}
    p_in = 1024
    for li, a in enumerate(layers):
        p_this = p_in
        p_in = -(-p_in // a["stride"])
        if li == 0:
            continue
        torch.manual_seed(0)
        conv = E.InterSO3Conv(a["dim_in"], a["dim_out"], 1, a["stride"], a["radius"], a["sigma"], a["n_neighbor"],
                              lazy_sample=True, kanchor=60).to(dev)
        g = torch.Generator().manual_seed(1)
        xyz = torch.randn(32, 3, p_this, generator=g)
        xyz = (xyz / xyz.norm(dim=1, keepdim=True)).to(dev)
        feats = torch.randn(32, a["dim_in"], p_this, 60, device=dev)
        x = E.SphericalPointCloud(xyz, feats, None)
        with torch.no_grad():
            idx, w, _, _ = conv(x)
            for _ in range(3):
                conv(x, idx, w) if a["stride"] == 1 else conv(x)
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(10):
                conv(x, idx, w) if a["stride"] == 1 else conv(x)
            e1.record()
            torch.cuda.synchronize()
        res["l%d_%dto%d_k%d" % (li, a["dim_in"], a["dim_out"], a["n_neighbor"])] = round(e0.elapsed_time(e1) / 10, 3)
    print(json.dumps(res))


def main():
    if "--one" in sys.argv:
        return one()
    out = {}
    for gather in (0, 1, 2):
        for sps in (0, 1):
            env = dict(os.environ, EPN_FU_GATHER=str(gather))
            if sps:
                env["EPN_FU_SPS"] = str(sps)
            r = subprocess.run([sys.executable, os.path.abspath(__file__), "--one"], env=env, capture_output=True, text=True)
            key = "gather%d_sps%s" % (gather, "auto" if not sps else sps)
            try:
                out[key] = json.loads(r.stdout.strip().splitlines()[-1])
            except Exception:
                out[key] = {"error": (r.stderr or r.stdout)[-300:]}
            print(key, out[key], flush=True)
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    json.dump(out, open(os.path.join(ROOT, "gpurun_out", "fused_sweep.json"), "w"), indent=1)


if __name__ == "__main__":
    main()
