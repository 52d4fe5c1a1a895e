#!/usr/bin/env python
"""Layer-by-layer parity report of the CUDA path against the oracle port (CPU, fp32 = the reference's op
chain, and fp64 = exact reference value), for the small golden backbone and for the full BASELINE
classification backbone (1024 pts, 60 anchors) at a small batch.  GPU box only (uses oracle/ as checker).

    python tools/parity_report.py [--full-batch 2] > profiles/parity_report.txt
"""
import argparse
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

from conftest import load_golden  # noqa: E402
from oracle import torch_port as TP  # noqa: E402
import epn_pointcloud_b200 as E  # noqa: E402
from epn_pointcloud_b200.blocks import SO3ConvBackbone, cls_backbone_params, preprocess_input  # noqa: E402


def rel(a, b):
    a, b = a.detach().double().cpu(), b.detach().double().cpu()
    return float((a - b).abs().max() / b.abs().max())


def frob(a, b):
    a, b = a.detach().double().cpu(), b.detach().double().cpu()
    return float((a - b).norm() / b.norm())


def port_chain(pc, layers, dtype):
    """Block outputs and weight grads of the oracle port in `dtype`."""
    leaves = []
    for prm, *_ in layers:
        for k in prm:
            prm[k] = prm[k].detach().to(dtype).requires_grad_(True)
        leaves.append(prm)
    b, n, _ = pc.shape
    xyz = pc.permute(0, 2, 1).contiguous()
    feats = torch.ones(b, 1, n, layers[0][3].shape[0], dtype=dtype)
    outs = []
    for prm, args, intra_idx, anchors, kernels in layers:
        xyz, feats = TP.separable_block(xyz, feats, prm, args, intra_idx, anchors, kernels)
        outs.append(feats)
    return outs, leaves


def gpu_chain(model, pc):
    outs = []
    x = preprocess_input(pc.cuda(), model.na_in, False)
    for block in model.backbone:
        for conv, param in zip(block.blocks, block.params):
            _, _, _, x = conv(x, None, None)
            outs.append(x.feats)
    return outs


def report(title, model, pc, r_seed):
    print("=" * 100)
    print(title)
    layers32 = TP.layers_from_module(model)
    layers64 = TP.layers_from_module(model)
    o32, l32 = port_chain(pc, layers32, torch.float32)
    o64, l64 = port_chain(pc, layers64, torch.float64)
    r = torch.randn(o32[-1].shape, generator=torch.Generator().manual_seed(r_seed))
    (o32[-1] * r).sum().backward()
    (o64[-1] * r.double()).sum().backward()
    for be in ("simt", "umma"):
        E.ops.set_gemm_backend(be)
        model.zero_grad()
        og = gpu_chain(model, pc)
        (og[-1] * r.cuda()).sum().backward()
        print("-- GEMM backend %s: forward, max|err|/max|ref|  (ours vs fp64 | reference-port fp32 vs fp64)" % be)
        for i, (a, b32, b64) in enumerate(zip(og, o32, o64)):
            print("   block %d out %-22s ours %.2e   ref-fp32 %.2e" % (i, tuple(a.shape), rel(a, b64), rel(b32, b64)))
        print("-- weight gradients, relative Frobenius error vs fp64")
        params = dict(model.named_parameters())
        layer = 0
        for bi, block in enumerate(model.backbone):
            for ci in range(len(block.blocks)):
                for sub, key in (("inter_conv", "inter_W"), ("intra_conv", "intra_W")):
                    n = "backbone.%d.blocks.%d.%s.conv.basic_conv.W" % (bi, ci, sub)
                    print("   %-58s ours %.2e   ref-fp32 %.2e" % (n, frob(params[n].grad, l64[layer][key].grad),
                                                                frob(l32[layer][key].grad, l64[layer][key].grad)))
                layer += 1
    E.ops.set_gemm_backend("umma")


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--full-batch", type=int, default=2)
    args = ap.parse_args()
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False

    g = load_golden("backbone_small")
    model = SO3ConvBackbone(g["params"], 60).cuda().train()
    model.load_state_dict(g.state_dict(), strict=True)
    report("small golden backbone (2 clouds x 128 pts, 3 separable blocks)", model, g["pc"], 23)

    torch.manual_seed(0)
    model = SO3ConvBackbone(cls_backbone_params(1024, 60), 60).cuda().train()
    gen = torch.Generator().manual_seed(2)
    pc = torch.randn(args.full_batch, 1024, 3, generator=gen)
    pc = pc / pc.norm(dim=2, keepdim=True)
    pc = pc - pc.mean(1, keepdim=True)
    pc = (pc / pc.norm(dim=2).amax(dim=1).view(-1, 1, 1)).contiguous()
    report("BASELINE cls backbone (%d clouds x 1024 pts, 60 anchors, 7 separable blocks)" % args.full_batch, model, pc, 24)


if __name__ == "__main__":
    main()
