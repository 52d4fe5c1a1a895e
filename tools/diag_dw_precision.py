"""Weight-gradient error at the BASELINE batch: the K dimension of the dW GEMM is the (cloud, point, anchor) index,
~1e6 long at 32 clouds, all accumulated in TMEM -- how far is dW from an fp64 evaluation (torch, on the GPU)?"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import epn_pointcloud_b200 as E  # noqa: E402

DEV = "cuda:0"


def rel(a, b):
    a, b = a.detach().double(), b.detach().double()
    return float((a - b).abs().max() / b.abs().max()), float((a - b).norm() / b.norm()), float((a * b).sum() / (b * b).sum() - 1.0)


def main():
    for B in (1, 8, 32):
        for (c, p) in ((64, 512), (256, 128)):
            torch.manual_seed(0)
            x = torch.randn(B, c, 1, p, 60, device=DEV)
            w = (torch.randn(c, c, device=DEV) / c ** 0.5).requires_grad_(True)
            y = E.modules._BasicConvFn.apply(x, w)
            r = torch.randn_like(y)
            (y * r).sum().backward()
            ref = torch.einsum("bopa,bcpa->oc", r.double(), x[:, :, 0].double())
            print("1x1 conv  B=%2d c=%3d p=%3d (K = %7d): dW max-rel %.2e rms-rel %.2e scale bias %+.2e" % ((B, c, p, B * p * 60) + rel(w.grad, ref)))
            intra = E.IntraSO3Conv(c, c).to(DEV)
            f = torch.randn(B, c, p, 60, device=DEV)
            z = intra(E.SphericalPointCloud(None, f, None)).feats
            r2 = torch.randn_like(z)
            (z * r2).sum().backward()
            idx = intra.intra_idx.long()                      # [60, 12]
            G = f.double()[:, :, :, idx]                      # [B, c, p, 60, 12]
            ref2 = torch.einsum("bopa,bcpak->ock", r2.double(), G).reshape(c, c * 12)
            print("intra     B=%2d c=%3d p=%3d (K = %7d): dW max-rel %.2e rms-rel %.2e scale bias %+.2e" % ((B, c, p, B * p * 60) + rel(intra.basic_conv.W.grad, ref2)))
            del G, ref2, ref


if __name__ == "__main__":
    main()
