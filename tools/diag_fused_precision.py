"""Where the fused inter conv loses precision: W = identity makes the kernel output the grouped tensor G itself
(out[o = c*24+k] = G[c,k]), which is compared with the fp64 oracle port for both operand formats, next to the op-level
grouping kernel (epn_inter_group_fwd_f32) and the layer-level errors with a random W."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import epn_pointcloud_b200 as E  # noqa: E402
from oracle import torch_port as TP  # noqa: E402

DEV = "cuda:0"


def rel(a, b):
    a, b = a.detach().cpu().double(), b.detach().cpu().double()
    return float((a - b).abs().max() / b.abs().max()), float((a - b).norm() / b.norm())


def main():
    for (c_in, p_in, stride, nn_, radius, sigma) in [(8, 512, 1, 16, 0.2828, 0.04), (8, 512, 2, 32, 0.4, 0.08)]:
        c_out = c_in * 24
        torch.manual_seed(1)
        conv = E.InterSO3Conv(c_in, c_out, 1, stride, radius, sigma, nn_, lazy_sample=True, kanchor=60).to(DEV)
        with torch.no_grad():
            conv.basic_conv.W.copy_(torch.eye(c_out))
        g = torch.Generator().manual_seed(5)
        xyz = torch.randn(1, 3, p_in, generator=g)
        xyz = (xyz / xyz.norm(dim=1, keepdim=True)).contiguous()
        feats = torch.randn(1, c_in, p_in, 60, generator=g)
        gx, idx, sidx, nx = TP.sample_and_query(xyz, stride, radius, nn_, True)
        for wd in (torch.float64, torch.float32):
            iw = TP.inter_weights(gx.to(wd), conv.anchors.cpu().to(wd), conv.kernels.cpu().to(wd), sigma).double()
            fsh = torch.cat((feats.double(), torch.zeros(1, c_in, 1, 60, dtype=torch.float64)), dim=2).contiguous()
            G = TP.inter_group(idx, iw, fsh)            # [b, c, k, p, a]
            if wd == torch.float64:
                G64 = G.reshape(1, c_out, G.shape[3], 60)
            else:
                print("  reference-style fp32 weights vs fp64: max-rel %.2e rms-rel %.2e" % rel(G.reshape(1, c_out, G.shape[3], 60), G64))
        x = E.SphericalPointCloud(xyz.to(DEV), feats.to(DEV), None)
        with torch.no_grad():
            for fmt in ("bf16", "f16"):
                with E.ops.forward_operands(fmt):
                    y = conv(x)[3].feats
                print("c_in %d K %d  fused kernel, W = I, %s operands: G max-rel %.2e rms-rel %.2e" % ((c_in, nn_, fmt) + rel(y, G64)))
            E.ops.set_fused_inter(False)
            y = conv(x)[3].feats
            print("  two-kernel route (bf16): max-rel %.2e rms-rel %.2e" % rel(y, G64))
            E.ops.set_fused_inter(True)
            ii = conv(x)[0]
            geom = (xyz.to(DEV), nx.to(DEV), conv.anchors, conv.kernels, sigma)
            Gk = E.ops.inter_group_fwd(feats.to(DEV), ii, None, geom)
            print("  op-level grouping kernel (fp32 out): max-rel %.2e rms-rel %.2e" % rel(Gk.reshape(1, c_out, -1, 60), G64))


if __name__ == "__main__":
    main()
