"""Generate tests/golden/*.npz by running the UNMODIFIED reference Python (vgtk + SPConvNets)
on CPU through oracle/ref_harness.py.  Build-container only:  python -m oracle.make_golden

TEST INFRASTRUCTURE ONLY.  The reference's three native index ops are served by the C oracle
(the reference kernels need a GPU; they are pinned separately on the GPU box through
oracle/_ref, see oracle/build_ref.py), everything else -- kernel weights, gather, einsum,
matmul, index_select, norms, skip branch, autograd -- is the reference's own code.

Every fixture stores inputs, the reference modules' state_dict and the reference outputs /
gradients, so the tests can (1) pin oracle/torch_port.py + oracle/epn_oracle.c on CPU and
(2) check the CUDA path on the GPU box where /root/reference does not exist.
"""
import json
import os
import types

import numpy as np
import torch

from oracle import ref_harness as H

OUT = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")


def sphere_points(b, n, seed):
    g = torch.Generator().manual_seed(seed)
    x = torch.randn(b, n, 3, generator=g)
    x = x / x.norm(dim=2, keepdim=True)
    x = x - x.mean(1, keepdim=True)
    return (x / x.norm(dim=2).amax(dim=1).view(b, 1, 1)).contiguous()


def npy(t):
    return t.detach().cpu().numpy()


def save(name, **arrays):
    path = os.path.join(OUT, name + ".npz")
    np.savez_compressed(path, **arrays)
    print("wrote %-28s %7.1f KB" % (name + ".npz", os.path.getsize(path) / 1024), {k: v.shape for k, v in arrays.items()})


def sd_arrays(module, prefix="sd."):
    return {prefix + k: npy(v) for k, v in module.state_dict().items()}


def main():
    os.makedirs(OUT, exist_ok=True)
    vgtk = H.load_reference()
    M = H.load_spconvnets()
    import vgtk.so3conv as sptk
    import vgtk.spconv as zptk
    import vgtk.so3conv.functional as RL

    # ---- constants (structural KATs of SURVEY.md section 4)
    save("so3_constants", anchors=np.ascontiguousarray(RL.get_anchors(60)), intra_idx=np.ascontiguousarray(RL.get_intra_idx()),
         kernels_r1=RL.get_sphereical_kernel_points_from_ply(0.7 * 0.4, 1),
         anchors20=np.ascontiguousarray(RL.get_anchors(20)), anchors40=np.ascontiguousarray(RL.get_anchors(40)))

    # ---- config 1: one 256-pt cloud, 20 anchors, InterSO3Conv (FPS + ball query + conv), occupancy feats
    torch.manual_seed(0)
    pc = sphere_points(1, 256, 0)
    conv = sptk.InterSO3Conv(1, 8, 1, 2, 0.4, 0.08, 16, lazy_sample=False, kanchor=20)
    x = M.preprocess_input(pc, 20, False)
    inter_idx, inter_w, sample_idx, y = conv(x)
    r = torch.randn(y.feats.shape, generator=torch.Generator().manual_seed(10))
    (y.feats * r).sum().backward()
    save("inter_a20_occupancy", pc=npy(pc), inter_idx=npy(inter_idx), sample_idx=npy(sample_idx),
         inter_w_p0_8=npy(inter_w[:, :8]), out=npy(y.feats), new_xyz=npy(y.xyz), r=npy(r),
         dW=npy(conv.basic_conv.W.grad), **sd_arrays(conv))

    # ---- InterSO3Conv, 60 anchors, real features, stride 1 and stride 2 (prefix sampling)
    for tag, stride, nn_, p_in in (("s1", 1, 16, 64), ("s2", 2, 32, 96)):
        torch.manual_seed(1)
        pc = sphere_points(2, p_in, 1)
        conv = sptk.InterSO3Conv(4, 8, 1, stride, 0.6, 0.18, nn_, lazy_sample=True, kanchor=60)
        feats = torch.randn(2, 4, p_in, 60, generator=torch.Generator().manual_seed(11)).requires_grad_(True)
        x = zptk.SphericalPointCloud(pc.permute(0, 2, 1).contiguous(), feats, None)
        inter_idx, inter_w, sample_idx, y = conv(x)
        r = torch.randn(y.feats.shape, generator=torch.Generator().manual_seed(12))
        (y.feats * r).sum().backward()
        save("inter_a60_" + tag, pc=npy(pc), feats=npy(feats), inter_idx=npy(inter_idx), sample_idx=npy(sample_idx),
             inter_w_p0_2=npy(inter_w[:, :2]), out=npy(y.feats), r=npy(r), dfeats=npy(feats.grad),
             dW=npy(conv.basic_conv.W.grad), **sd_arrays(conv))

    # ---- grouping stages alone (op surface): inter_zpconv_grouping_naive, intra_so3conv_grouping
    torch.manual_seed(2)
    pc = sphere_points(1, 48, 2)
    xyz = pc.permute(0, 2, 1).contiguous()
    anchors = torch.from_numpy(RL.get_anchors(60))
    kernels = torch.from_numpy(RL.get_sphereical_kernel_points_from_ply(0.7 * 0.5, 1))
    gxyz, ball_idx, sidx, sxyz = zptk.functional.inter_zpconv_grouping_ball(xyz, 1, 0.5, 12, True)
    w = RL.inter_so3conv_grouping_anchor(gxyz, anchors, kernels, 0.125)
    feats = torch.randn(1, 3, 48, 60, generator=torch.Generator().manual_seed(13)).requires_grad_(True)
    g = zptk.functional.inter_zpconv_grouping_naive(ball_idx, w, zptk.functional.add_shadow_feature(feats))
    r = torch.randn(g.shape, generator=torch.Generator().manual_seed(14))
    (g * r).sum().backward()
    save("inter_group", pc=npy(pc), anchors=npy(anchors), kernels=npy(kernels), sigma=np.float32(0.125),
         ball_idx=npy(ball_idx), grouped_xyz=npy(gxyz), inter_w=npy(w), feats=npy(feats), grouped=npy(g), r=npy(r),
         dfeats=npy(feats.grad))

    torch.manual_seed(3)
    intra_idx = torch.from_numpy(RL.get_intra_idx()).long()
    feats = torch.randn(2, 3, 10, 60, generator=torch.Generator().manual_seed(15)).requires_grad_(True)
    g = RL.intra_so3conv_grouping(intra_idx, feats)
    r = torch.randn(g.shape, generator=torch.Generator().manual_seed(16))
    (g * r).sum().backward()
    save("intra_group", feats=npy(feats), grouped=npy(g), r=npy(r), dfeats=npy(feats.grad))

    # ---- IntraSO3Conv
    torch.manual_seed(4)
    conv = sptk.IntraSO3Conv(4, 8)
    feats = torch.randn(2, 4, 32, 60, generator=torch.Generator().manual_seed(17)).requires_grad_(True)
    y = conv(zptk.SphericalPointCloud(torch.zeros(2, 3, 32), feats, None))
    r = torch.randn(y.feats.shape, generator=torch.Generator().manual_seed(18))
    (y.feats * r).sum().backward()
    save("intra_a60", feats=npy(feats), out=npy(y.feats), r=npy(r), dfeats=npy(feats.grad),
         dW=npy(conv.basic_conv.W.grad), **sd_arrays(conv))

    # ---- BasicSO3Conv on a grouped tensor
    torch.manual_seed(5)
    conv = sptk.BasicSO3Conv(3, 5, 24)
    xg = torch.randn(2, 3, 24, 7, 60, generator=torch.Generator().manual_seed(19)).requires_grad_(True)
    y = conv(xg)
    r = torch.randn(y.shape, generator=torch.Generator().manual_seed(20))
    (y * r).sum().backward()
    save("basic_conv", x=npy(xg), out=npy(y), r=npy(r), dx=npy(xg.grad), dW=npy(conv.W.grad), **sd_arrays(conv))

    # ---- one SeparableSO3ConvBlock (inter + BN + intra + IN + strided skip), training mode
    torch.manual_seed(6)
    args = {"dim_in": 4, "dim_out": 8, "kernel_size": 1, "stride": 2, "radius": 0.6, "sigma": 0.18, "n_neighbor": 16,
            "lazy_sample": True, "dropout_rate": 0.0, "multiplier": 2, "activation": "leaky_relu", "pooling": None,
            "kanchor": 60, "norm": "BatchNorm2d"}
    blk = M.SeparableSO3ConvBlock(dict(args)).train()
    pc = sphere_points(2, 64, 6)
    feats = torch.randn(2, 4, 64, 60, generator=torch.Generator().manual_seed(21)).requires_grad_(True)
    sd0 = sd_arrays(blk)  # before the forward updates the BN running stats
    _, _, _, y = blk(zptk.SphericalPointCloud(pc.permute(0, 2, 1).contiguous(), feats, None), None, None)
    r = torch.randn(y.feats.shape, generator=torch.Generator().manual_seed(22))
    (y.feats * r).sum().backward()
    grads = {"grad." + k: npy(p.grad) for k, p in blk.named_parameters()}
    save("separable_block", pc=npy(pc), feats=npy(feats), out=npy(y.feats), r=npy(r), dfeats=npy(feats.grad),
         args=np.array(json.dumps(args)), **sd0, **grads)

    # ---- small classification-style backbone (2 blocks, 3 separable layers) end to end
    from epn_pointcloud_b200.blocks import cls_backbone_params
    bp = cls_backbone_params(input_num=128, mlps=((8, 8), (16,)), strides=(2, 2), initial_radius_ratio=0.4,
                             sampling_ratio=0.8)
    torch.manual_seed(7)
    backbone = torch.nn.ModuleList([M.BasicSO3ConvBlock(b) for b in bp]).train()
    pc = sphere_points(2, 128, 7)
    sd0 = {"sd.backbone." + k: npy(v) for k, v in backbone.state_dict().items()}
    x = M.preprocess_input(pc, 60, False)
    for blk in backbone:
        x = blk(x)
    r = torch.randn(x.feats.shape, generator=torch.Generator().manual_seed(23))
    (x.feats * r).sum().backward()
    grads = {"grad.backbone." + k: npy(p.grad) for k, p in backbone.named_parameters()}
    save("backbone_small", pc=npy(pc), out=npy(x.feats), out_xyz=npy(x.xyz), r=npy(r),
         params=np.array(json.dumps(bp)), **sd0, **grads)

    # ---- classification head (ClsOutBlockPointnet + PointnetSO3Conv), training mode
    torch.manual_seed(8)
    hp = {"dim_in": 16, "mlp": [24], "fc": [64], "k": 40, "pooling": "max", "temperature": 3.0, "kanchor": 60}
    head = M.ClsOutBlockPointnet(dict(hp)).train()
    pc = sphere_points(3, 32, 8)
    feats = torch.randn(3, 16, 32, 60, generator=torch.Generator().manual_seed(24)).requires_grad_(True)
    sd0 = sd_arrays(head)
    logits, hfeat = head(zptk.SphericalPointCloud(pc.permute(0, 2, 1).contiguous(), feats, None))
    r = torch.randn(logits.shape, generator=torch.Generator().manual_seed(25))
    (logits * r).sum().backward()
    grads = {"grad." + k: npy(p.grad) for k, p in head.named_parameters()}
    save("cls_head", pc=npy(pc), feats=npy(feats), logits=npy(logits), hfeat=npy(hfeat), r=npy(r),
         dfeats=npy(feats.grad), params=np.array(json.dumps(hp)), **sd0, **grads)

    # ---- the reference's own layer arithmetic for the three shipped models at the BASELINE sizes
    def opt_for(input_num, kanchor=60):
        o = types.SimpleNamespace()
        o.device = "cpu"
        o.model = types.SimpleNamespace(input_num=input_num, dropout_rate=0.0, kpconv=False, kanchor=kanchor,
                                        flag="max", search_radius=0.4)
        o.train_loss = types.SimpleNamespace(temperature=3.0)
        return o

    import importlib
    import contextlib
    import io
    out = {}
    for name, mod, n in (("cls", "SPConvNets.models.cls_so3net_pn", 1024),):
        m = importlib.import_module(mod)
        path = os.path.join(OUT, "_tmp_params.json")
        with contextlib.redirect_stdout(io.StringIO()):
            model = m.build_model(opt_for(n), to_file=path)
        out[name] = json.load(open(path))["backbone"]
        out[name + "_n_params"] = sum(p.numel() for p in model.parameters())
        out[name + "_state_keys"] = sorted(k for k in model.state_dict().keys() if k.startswith("backbone."))
        out[name + "_state_shapes_all"] = {k: list(v.shape) for k, v in model.state_dict().items()}
        out[name + "_outblock"] = json.load(open(path))["outblock"]
        os.remove(path)
    json.dump(out, open(os.path.join(OUT, "model_params.json"), "w"), indent=0)
    print("wrote model_params.json", out["cls_n_params"])


if __name__ == "__main__":
    main()
