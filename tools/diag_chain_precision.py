"""Where the whole-network forward error comes from: the BASELINE cls backbone at a small batch under no_grad,
(1) chained block outputs vs the fp64 oracle port (next to the fp32 port = the reference's own fp32 chain), and
(2) every STAGE of every block in isolation -- the stage's input is the fp64 port's value rounded to fp32, so each line
is the error that stage alone adds."""
import os
import sys

import torch
import torch.nn.functional as F

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import epn_pointcloud_b200 as E  # noqa: E402
from epn_pointcloud_b200 import blocks  # noqa: E402
from epn_pointcloud_b200.blocks import SO3ConvBackbone, cls_backbone_params, norm_act, preprocess_input  # noqa: E402
from epn_pointcloud_b200 import modules as sptk  # noqa: E402
from oracle import torch_port as TP  # noqa: E402

DEV = "cuda:0"


def rel(a, b):
    a, b = a.detach().double().cpu(), b.detach().double().cpu()
    return float((a - b).abs().max() / b.abs().max()), float((a - b).norm() / b.norm())


def main():
    B = int(sys.argv[1]) if len(sys.argv) > 1 else 2
    torch.backends.cudnn.allow_tf32 = torch.backends.cuda.matmul.allow_tf32 = False
    torch.manual_seed(0)
    model = SO3ConvBackbone(cls_backbone_params(1024, 60), 60).to(DEV).train()
    gen = torch.Generator().manual_seed(2)
    pc = torch.randn(B, 1024, 3, generator=gen)
    pc = pc / pc.norm(dim=2, keepdim=True)
    pc = pc - pc.mean(1, keepdim=True)
    pc = (pc / pc.norm(dim=2).amax(dim=1).view(-1, 1, 1)).contiguous()
    layers = TP.layers_from_module(model)
    convs = [c for blk in model.backbone for c in blk.blocks]

    def port(dtype):
        xyz = pc.permute(0, 2, 1).contiguous()
        feats = torch.ones(B, 1, 1024, 60, dtype=dtype)
        stages = []
        for prm, args, intra_idx, anchors, kernels in layers:
            prm = {k: v.to(dtype) for k, v in prm.items()}
            norm = args.get("norm")
            skip = feats
            _, _, sidx, nxyz, x0 = TP.inter_so3conv(xyz, feats, prm["inter_W"], anchors, kernels, args["stride"], args["n_neighbor"],
                                                    args["radius"], args["sigma"], args["lazy_sample"])
            x1 = F.leaky_relu(TP._norm(x0, norm, prm.get("inter_bn_w"), prm.get("inter_bn_b")))
            x2 = TP.intra_so3conv(x1, prm["intra_W"], intra_idx)
            x3 = F.leaky_relu(TP._norm(x2, None))
            if args["stride"] > 1:
                b, c, _, a = skip.shape
                index = sidx.long().view(b, 1, -1, 1).expand(b, c, sidx.shape[1], a)
                skip = torch.gather(skip, 2, index)
            s0 = F.conv2d(skip, prm["skip_w"], prm["skip_b"])
            s1 = F.leaky_relu(TP._norm(s0, norm, prm.get("bn_w"), prm.get("bn_b")))
            out = x3 + s1
            stages.append(dict(xyz=xyz, fin=feats, nxyz=nxyz, x0=x0, x1=x1, x2=x2, x3=x3, skip=skip, s0=s0, out=out))
            xyz, feats = nxyz, out
        return stages

    st64, st32 = port(torch.float64), port(torch.float32)
    for fmt in ("f16", "bf16"):
        blocks.set_inference_operands(fmt)
        print("=== no_grad forward, inference operands %s: chained block outputs, max-rel / rms-rel vs fp64 (fp32 port in brackets)" % fmt)
        with torch.no_grad():
            x = preprocess_input(pc.to(DEV), 60, False)
            for i, conv in enumerate(convs):
                _, _, _, x = conv(x, None, None)
                print("  block %d  ours %.2e / %.2e   [fp32 port %.2e / %.2e]" % ((i,) + rel(x.feats, st64[i]["out"]) + rel(st32[i]["out"], st64[i]["out"])))
    blocks.set_inference_operands("f16")
    print("=== stages in isolation (input = fp64 port value rounded to fp32), f16 operands: max-rel / rms-rel vs fp64 [fp32 port stage error]")
    with torch.no_grad():
        for i, conv in enumerate(convs):
            s = st64[i]
            fin = blocks._mark_unit(s["fin"].float().to(DEV))
            xin = sptk.SphericalPointCloud(s["xyz"].to(DEV), fin, None) if i > 0 else preprocess_input(pc.to(DEV), 60, False)
            with blocks.fwd_operands(fin, occupancy=(i == 0)):
                _, _, _, y = conv.inter_conv.conv(xin, None, None)
            e0 = rel(y.feats, s["x0"])
            e1 = rel(norm_act(conv.inter_conv.norm, s["x0"].float().to(DEV), F.leaky_relu), s["x1"])
            zin = blocks._mark_unit(s["x1"].float().to(DEV))
            with blocks.fwd_operands(zin):
                z = conv.intra_conv.conv(sptk.SphericalPointCloud(None, zin, None)).feats
            e2 = rel(z, s["x2"])
            e3 = rel(norm_act(conv.intra_conv.norm, s["x2"].float().to(DEV), F.leaky_relu), s["x3"])
            w = conv.skip_conv.weight.view(conv.skip_conv.out_channels, conv.skip_conv.in_channels)
            with blocks.fwd_operands(occupancy=True):
                sk = sptk._BasicConvFn.apply(s["skip"].float().to(DEV).unsqueeze(2).contiguous(), w)
            e4 = rel(sk + conv.skip_conv.bias.view(1, -1, 1, 1), s["s0"])
            e5 = rel(norm_act(conv.norm, sk, F.leaky_relu, residual=s["x3"].float().to(DEV), bias=conv.skip_conv.bias), s["out"])
            print("  block %d: inter %.1e/%.1e  norm %.1e/%.1e  intra %.1e/%.1e  norm %.1e/%.1e  skip conv %.1e/%.1e  skip norm+add %.1e/%.1e"
                  % ((i,) + e0 + e1 + e2 + e3 + e4 + e5))


if __name__ == "__main__":
    main()
