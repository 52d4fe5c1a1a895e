"""CPU: pin the oracle (oracle/epn_oracle.c + oracle/torch_port.py) against fixtures produced by the
reference's own Python (oracle/make_golden.py -> tests/golden/*.npz)."""
import json
import os

import numpy as np
import torch

from conftest import GOLDEN, load_golden, rel_err
from oracle import epn_oracle as O
from oracle import torch_port as TP

TOL = 2e-5  # fp32 op-order noise between the reference chain and the port / double-accumulating C oracle


def test_constants_structural_kats():
    g = load_golden("so3_constants")
    R = g["anchors"].double()
    assert R.shape == (60, 3, 3)
    eye = torch.eye(3, dtype=torch.float64)
    assert (R @ R.transpose(1, 2) - eye).abs().max() < 1e-5
    assert (torch.linalg.det(R) - 1).abs().max() < 1e-5
    assert (R[29] - eye).abs().max() == 0
    prod = torch.einsum("aij,bjk->abik", R, R).reshape(3600, 9)
    d = torch.cdist(prod, R.reshape(60, 9))
    assert d.min(dim=1).values.max() < 1e-4  # closed under multiplication
    ii = g["intra_idx"]
    assert ii.shape == (60, 12)
    assert ii[0].tolist() == [10, 33, 13, 15, 56, 30, 59, 8, 44, 0, 45, 26]
    assert ii[29].tolist() == [44, 31, 32, 13, 42, 43, 30, 14, 12, 29, 28, 27]
    assert ii[59].tolist() == [45, 43, 48, 26, 12, 40, 0, 32, 5, 59, 10, 15]
    for k in range(12):
        assert sorted(ii[:, k].tolist()) == list(range(60))
    # right-multiplication structure: intra_idx[a,k] = index_of(R_a R_0^T R_{intra_idx[0,k]})
    for a in (1, 7, 42):
        for k in range(12):
            M = R[a] @ R[0].T @ R[ii[0, k]]
            assert int((R - M).abs().amax(dim=(1, 2)).argmin()) == int(ii[a, k])


def test_product_constants_match_reference():
    from epn_pointcloud_b200 import functional as L
    g = load_golden("so3_constants")
    assert np.array_equal(L.get_anchors(60), g["anchors"].numpy())
    assert np.array_equal(L.get_anchors(20), g["anchors20"].numpy())
    assert np.array_equal(L.get_anchors(40), g["anchors40"].numpy())
    assert np.array_equal(L.get_intra_idx(), g["intra_idx"].numpy())
    assert np.array_equal(L.get_sphereical_kernel_points_from_ply(0.7 * 0.4, 1), g["kernels_r1"].numpy())
    k = g["kernels_r1"]
    assert k.shape == (24, 3) and abs(float(k.norm(dim=1).max()) - 0.28) < 1e-6 and float(k[0].norm()) == 0.0


def test_model_param_arithmetic_and_state_keys():
    from epn_pointcloud_b200.blocks import SO3ConvBackbone, cls_backbone_params
    ref = json.load(open(os.path.join(GOLDEN, "model_params.json")))
    mine = json.loads(json.dumps(cls_backbone_params(1024, 60)))
    assert mine == ref["cls"]
    model = SO3ConvBackbone(cls_backbone_params(1024, 60), 60)
    assert sorted(model.state_dict().keys()) == ref["cls_state_keys"]


def test_full_cls_model_matches_reference_structure():
    """ClsSO3ConvModel (backbone + ClsOutBlockPointnet head): same parameter/buffer names, shapes and count
    as the reference's build_model output (7,814,632 parameters), so its checkpoints load key for key."""
    from epn_pointcloud_b200.heads import ClsSO3ConvModel, cls_model_params
    ref = json.load(open(os.path.join(GOLDEN, "model_params.json")))
    params = json.loads(json.dumps(cls_model_params(1024, 60)))
    assert params["outblock"] == ref["cls_outblock"] and params["backbone"] == ref["cls"]
    model = ClsSO3ConvModel(cls_model_params(1024, 60))
    mine = {k: list(v.shape) for k, v in model.state_dict().items()}
    assert mine == ref["cls_state_shapes_all"]
    assert sum(p.numel() for p in model.parameters()) == ref["cls_n_params"]


def test_cls_head_port():
    from epn_pointcloud_b200.heads import ClsOutBlockPointnet
    g = load_golden("cls_head")
    head = ClsOutBlockPointnet(dict(g["params"]))
    head.load_state_dict(g.state_dict(), strict=True)
    hp = TP.head_from_module(head)
    logits, hfeat = TP.cls_head(g["pc"].permute(0, 2, 1).contiguous(), g["feats"], hp)
    assert rel_err(logits, g["logits"]) < TOL and rel_err(hfeat, g["hfeat"]) < TOL


def test_inter_a20_occupancy():
    g = load_golden("inter_a20_occupancy")
    sd = g.state_dict()
    xyz = g["pc"].permute(0, 2, 1).contiguous()
    feats = torch.ones(1, 1, 256, 20)
    W = sd["basic_conv.W"].clone().requires_grad_(True)
    idx, w, sidx, new_xyz, out = TP.inter_so3conv(xyz, feats, W, sd["anchors"], sd["kernels"], 2, 16, 0.4, 0.08, False)
    assert torch.equal(idx, g["inter_idx"]) and torch.equal(sidx, g["sample_idx"])
    assert torch.equal(new_xyz, g["new_xyz"])
    assert rel_err(w[:, :8], g["inter_w_p0_8"]) < TOL
    assert rel_err(out, g["out"]) < TOL
    (out * g["r"]).sum().backward()
    assert rel_err(W.grad, g["dW"]) < TOL
    # C oracle, composed
    _, w2, _, out2, _ = O.inter_so3conv(xyz, feats, sd["basic_conv.W"], sd["anchors"], sd["kernels"], 2, 16, 0.4, 0.08, False)
    assert rel_err(w2[:, :8], g["inter_w_p0_8"]) < TOL and rel_err(out2, g["out"]) < TOL


def _inter_case(name, stride, nn_):
    g = load_golden(name)
    sd = g.state_dict()
    xyz = g["pc"].permute(0, 2, 1).contiguous()
    feats = g["feats"].clone().requires_grad_(True)
    W = sd["basic_conv.W"].clone().requires_grad_(True)
    idx, w, sidx, _, out = TP.inter_so3conv(xyz, feats, W, sd["anchors"], sd["kernels"], stride, nn_, 0.6, 0.18, True)
    assert torch.equal(idx, g["inter_idx"]) and torch.equal(sidx, g["sample_idx"])
    assert rel_err(w[:, :2], g["inter_w_p0_2"]) < TOL
    assert rel_err(out, g["out"]) < TOL
    (out * g["r"]).sum().backward()
    assert rel_err(feats.grad, g["dfeats"]) < TOL and rel_err(W.grad, g["dW"]) < TOL
    _, _, _, out2, _ = O.inter_so3conv(xyz, g["feats"], sd["basic_conv.W"], sd["anchors"], sd["kernels"], stride, nn_, 0.6, 0.18, True)
    assert rel_err(out2, g["out"]) < TOL


def test_inter_a60_stride1():
    _inter_case("inter_a60_s1", 1, 16)


def test_inter_a60_stride2():
    _inter_case("inter_a60_s2", 2, 32)


def test_inter_group_stages():
    g = load_golden("inter_group")
    xyz = g["pc"].permute(0, 2, 1).contiguous()
    sigma = float(g["sigma"])
    idx = O.ball_query(xyz, xyz, 0.5, 12)
    assert torch.equal(idx, g["ball_idx"])
    w = O.inter_weights(xyz, xyz, idx, g["anchors"], g["kernels"], sigma)
    assert rel_err(w, g["inter_w"]) < TOL
    assert rel_err(TP.inter_weights(g["grouped_xyz"], g["anchors"], g["kernels"], sigma), g["inter_w"]) < TOL
    assert rel_err(O.inter_group_fwd(idx, g["inter_w"], g["feats"]), g["grouped"]) < TOL
    assert rel_err(O.inter_group_bwd(idx, g["inter_w"], g["r"], 48), g["dfeats"]) < TOL
    assert rel_err(TP.inter_group(idx, g["inter_w"], g["feats"]), g["grouped"]) < TOL


def test_intra_group_stages():
    g = load_golden("intra_group")
    ii = load_golden("so3_constants")["intra_idx"]
    assert torch.equal(O.intra_group_fwd(ii, g["feats"]), g["grouped"])
    assert torch.equal(TP.intra_group(ii, g["feats"]), g["grouped"])
    assert rel_err(O.intra_group_bwd(ii, g["r"]), g["dfeats"]) < TOL


def test_intra_a60():
    g = load_golden("intra_a60")
    sd = g.state_dict()
    feats = g["feats"].clone().requires_grad_(True)
    W = sd["basic_conv.W"].clone().requires_grad_(True)
    out = TP.intra_so3conv(feats, W, sd["intra_idx"])
    assert rel_err(out, g["out"]) < TOL
    (out * g["r"]).sum().backward()
    assert rel_err(feats.grad, g["dfeats"]) < TOL and rel_err(W.grad, g["dW"]) < TOL
    assert rel_err(O.intra_so3conv(g["feats"], sd["basic_conv.W"], sd["intra_idx"]), g["out"]) < TOL


def test_basic_conv():
    g = load_golden("basic_conv")
    W = g.state_dict()["W"]
    assert rel_err(O.basic_conv(g["x"], W), g["out"]) < TOL
    assert rel_err(TP.basic_conv(g["x"], W), g["out"]) < TOL


def test_separable_block_port():
    g = load_golden("separable_block")
    sd = g.state_dict()
    prm = {"inter_W": sd["inter_conv.conv.basic_conv.W"], "inter_bn_w": sd["inter_conv.norm.weight"],
           "inter_bn_b": sd["inter_conv.norm.bias"], "intra_W": sd["intra_conv.conv.basic_conv.W"],
           "skip_w": sd["skip_conv.weight"], "skip_b": sd["skip_conv.bias"], "bn_w": sd["norm.weight"],
           "bn_b": sd["norm.bias"]}
    prm = {k: v.clone().requires_grad_(True) for k, v in prm.items()}
    feats = g["feats"].clone().requires_grad_(True)
    xyz = g["pc"].permute(0, 2, 1).contiguous()
    _, out = TP.separable_block(xyz, feats, prm, g["args"], sd["intra_conv.conv.intra_idx"],
                                sd["inter_conv.conv.anchors"], sd["inter_conv.conv.kernels"])
    assert rel_err(out, g["out"]) < 1e-4
    (out * g["r"]).sum().backward()
    gr = g.grads()
    assert rel_err(feats.grad, g["dfeats"]) < 1e-3
    assert rel_err(prm["inter_W"].grad, gr["inter_conv.conv.basic_conv.W"]) < 1e-3
    assert rel_err(prm["intra_W"].grad, gr["intra_conv.conv.basic_conv.W"]) < 1e-3
    assert rel_err(prm["skip_w"].grad, gr["skip_conv.weight"]) < 1e-3


def test_backbone_small_port():
    from epn_pointcloud_b200.blocks import SO3ConvBackbone
    g = load_golden("backbone_small")
    model = SO3ConvBackbone(g["params"], 60)
    missing = model.load_state_dict(g.state_dict(), strict=True)  # reference keys load one to one
    assert not missing.missing_keys and not missing.unexpected_keys
    layers = TP.layers_from_module(model)
    xyz, out = TP.backbone_forward(g["pc"], layers)
    assert torch.equal(xyz, g["out_xyz"])
    assert rel_err(out, g["out"]) < 1e-4


def test_inv_and_reg_model_arithmetic_and_state_dicts_match_reference():
    """The layer arithmetic (radius / sigma / neighbour count / stride per layer) and every state-dict key and
    shape of the full-size 3DMatch (2048 pts) and rotation (1024 pts) models equal what the reference's own
    build_model produced (tests/golden/model_params_inv_reg.json, oracle/make_golden_models.py)."""
    import json
    import os
    from epn_pointcloud_b200 import heads
    g = json.load(open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "model_params_inv_reg.json")))
    for name, params, cls in (("inv", heads.inv_model_params(2048, 60), heads.InvSO3ConvModel),
                              ("reg", heads.reg_model_params(1024, 60), heads.RegSO3ConvModel)):
        assert json.loads(json.dumps(params["backbone"])) == g[name]
        assert params["outblock"] == g[name + "_outblock"]
        model = cls(params)
        shapes = {k: list(v.shape) for k, v in model.state_dict().items()}
        ref = g[name + "_state_shapes_all"]
        assert {k: v for k, v in shapes.items() if not k.endswith("_intra_idx32")} == ref
        assert sum(p.numel() for p in model.parameters()) == g[name + "_n_params"]


def _strip(sd, prefix):
    return {k[len(prefix):]: v.float() for k, v in sd.items() if k.startswith(prefix)}


def test_port_matches_reference_inv_model():
    """oracle/torch_port.py (backbone + InvOutBlockMVD restatement) against the reference's own 3DMatch model output
    (tests/golden/inv_model_small.npz, made by oracle/make_golden_models.py): pins the CPU oracle for row 8 f2."""
    import torch
    from conftest import load_golden, rel_err
    from epn_pointcloud_b200.heads import InvSO3ConvModel
    from oracle import torch_port as TP
    g = load_golden("inv_model_small")
    model = InvSO3ConvModel(g["params"])
    model.load_state_dict(g.state_dict(), strict=True)
    with torch.no_grad():
        xyz, feats = TP.backbone_forward(g["pc"], TP.layers_from_module(model))
        desc, attn = TP.inv_head(xyz, feats, TP.inv_head_from_state(_strip(model.state_dict(), "outblock.")))
    assert rel_err(desc, g["desc"]) < 1e-5 and rel_err(attn, g["attn"]) < 1e-5


def test_port_matches_reference_reg_model():
    """Same for the relative-rotation model (shared backbone over both clouds of a pair + RelSO3OutBlockR)."""
    import torch
    from conftest import load_golden, rel_err
    from epn_pointcloud_b200.heads import RegSO3ConvModel
    from oracle import torch_port as TP
    g = load_golden("reg_model_small")
    model = RegSO3ConvModel(g["params"])
    model.load_state_dict(g.state_dict(), strict=True)
    x = torch.cat((g["pairs"][:, 0], g["pairs"][:, 1]), dim=0)
    with torch.no_grad():
        xyz, feats = TP.backbone_forward(x, TP.layers_from_module(model))
        f1, f2 = torch.chunk(feats, 2, dim=0)
        x1, x2 = torch.chunk(xyz, 2, dim=0)
        hp = TP.rel_head_from_state(_strip(model.state_dict(), "outblock."), g["params"]["outblock"]["temperature"])
        conf, quats = TP.rel_head(f1, f2, x1, x2, hp)
    assert rel_err(conf, g["conf"]) < 1e-5 and rel_err(quats, g["quats"]) < 1e-5
