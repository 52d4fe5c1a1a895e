"""CPU restatement ("port") of the reference's PyTorch op chains for the SPConv hot path.

TEST INFRASTRUCTURE ONLY -- the parity checker and the CPU baseline.  Only tests/,
__graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs import this;
the product package `epn_pointcloud_b200` never does.

It keeps the reference's op chain (torch.gather -> broadcast weights -> einsum -> matmul;
index_select -> permute -> matmul; norm -> leaky_relu -> skip) so that timing it on host
cores measures what the reference's own CPU path costs, and its numerics are fp32 like
the reference.  The three native index ops come from the C oracle (oracle/epn_oracle.c).
Pinned against the real reference by oracle/make_golden.py (tests/golden/*.npz).

Each function cites the reference file:line it follows (nintendops/EPN_PointCloud @ b625483).
"""
import math

import torch
import torch.nn.functional as F

from oracle import epn_oracle as O


class _RefCudaOps:
    """Index ops for CUDA tensors: the REFERENCE's own CUDA kernels (oracle/_ref, built by oracle/build_ref.py
    from vgtk/vgtk/cuda/*), so that running this port on a GPU is the reference's GPU path -- its PyTorch op
    chain plus its native extensions -- and not anything of the product library."""
    _mods = {}

    @classmethod
    def _mod(cls, ext):
        if ext not in cls._mods:
            from oracle import build_ref
            m = build_ref.load_ref(ext)
            if m is None:
                raise RuntimeError("oracle/_ref/vgtk_ref_%s.so is missing (python -m oracle.build_ref)" % ext)
            cls._mods[ext] = m
        return cls._mods[ext]

    @classmethod
    def furthest_point_sampling(cls, xyz, m):
        return cls._mod("grouping").furthest_point_sampling(xyz.contiguous(), m)

    @classmethod
    def ball_query(cls, new_xyz, xyz, radius, nsample):
        return cls._mod("grouping").ball_query(new_xyz.contiguous(), xyz.contiguous(), radius, nsample)

    @classmethod
    def gather_points_forward(cls, points, idx):
        return cls._mod("gathering").gather_points_forward(points.contiguous(), idx.contiguous())


def _ops(t):
    return _RefCudaOps if t.is_cuda else O


# ------------------------------------------------------------------ sampling
def sample_and_query(xyz, stride, radius, n_neighbor, lazy_sample):
    """vgtk/vgtk/spconv/functional.py:412-421 + vgtk/vgtk/pc/sample.py:46-77.
    xyz [b,3,p_in] -> (grouped_xyz - centre [b,3,p,nn], ball_idx [b,p,nn], sample_idx [b,p], new_xyz [b,3,p])"""
    b, _, p_in = xyz.shape
    n_sample = math.ceil(p_in / stride)
    if p_in == n_sample or lazy_sample:
        sample_idx = torch.arange(n_sample, dtype=torch.int32, device=xyz.device).view(1, -1).expand(b, -1).contiguous()
    else:
        sample_idx = _ops(xyz).furthest_point_sampling(xyz, n_sample)
    new_xyz = _ops(xyz).gather_points_forward(xyz, sample_idx)
    ball_idx = _ops(xyz).ball_query(new_xyz, xyz, radius, n_neighbor)
    grouped = _ops(xyz).gather_points_forward(xyz, ball_idx.view(b, -1)).view(b, 3, n_sample, n_neighbor)
    return grouped - new_xyz.unsqueeze(3), ball_idx, sample_idx, new_xyz


# ------------------------------------------------------------ kernel weights
def inter_weights(grouped_xyz, anchors, kernels, sigma):
    """vgtk/vgtk/so3conv/functional.py:180-218: rotate kernel points by every anchor, squared
    distance to every neighbour offset, linear falloff clipped at 0 -> [b,p,na,ks,nn]."""
    rk = torch.matmul(anchors, kernels.t())                 # [na,3,ks]
    rk = rk.permute(1, 0, 2).contiguous()                   # [3,na,ks]
    diff = grouped_xyz[:, :, :, None, None, :] - rk[None, :, None, :, :, None]   # [b,3,p,na,ks,nn]
    dist2 = (diff ** 2).sum(dim=1)
    return F.relu(1.0 - dist2 / sigma)


# ------------------------------------------------------------ feature grouping
def inter_group(inter_idx, inter_w, feats):
    """vgtk/vgtk/spconv/functional.py:361-390: expanded-index gather of the neighbours' feature rows,
    then the per-anchor spatial contraction over the nn neighbours -> [b,c,ks,p,na]."""
    b, p, nn_ = inter_idx.shape
    _, c, q, a = feats.shape
    index = inter_idx.long().view(b, 1, p * nn_, 1).expand(b, c, p * nn_, a)
    g = torch.gather(feats, 2, index).view(b, c, p, nn_, a)
    return torch.einsum("bcpna,bpakn->bckpa", g, inter_w).contiguous()


def intra_group(intra_idx, feats):
    """vgtk/vgtk/so3conv/functional.py:221-268 -> [b,c,kn,p,na]"""
    b, c, p, na = feats.shape
    kn = intra_idx.shape[1]
    g = feats.index_select(3, intra_idx.long().reshape(-1)).view(b, c, p, na, kn)
    return g.permute(0, 1, 4, 2, 3).contiguous()


def basic_conv(x, W):
    """vgtk/vgtk/so3conv/modules.py:48-55"""
    b, c, ks, p, a = x.shape
    return torch.matmul(W, x.view(b, c * ks, p * a)).view(b, W.shape[0], p, a)


# ----------------------------------------------------------------- conv layers
def inter_so3conv(xyz, feats, W, anchors, kernels, stride, n_neighbor, radius, sigma, lazy_sample=True):
    """InterSO3Conv.forward: vgtk/vgtk/so3conv/modules.py:157-174 -> so3conv/functional.py:118-178
    (incl. the zero shadow row appended at :174, spconv/functional.py:91-95).
    Returns (inter_idx, inter_w, sample_idx, new_xyz, out)."""
    grouped_xyz, inter_idx, sample_idx, new_xyz = sample_and_query(xyz, stride, radius, n_neighbor, lazy_sample)
    inter_w = inter_weights(grouped_xyz.to(feats.dtype), anchors.to(feats.dtype), kernels.to(feats.dtype), sigma)
    b, c, _, a = feats.shape
    feats_sh = torch.cat((feats, torch.zeros(b, c, 1, a, dtype=feats.dtype, device=feats.device)), dim=2).contiguous()
    grouped = inter_group(inter_idx, inter_w, feats_sh)
    return inter_idx, inter_w, sample_idx, new_xyz, basic_conv(grouped, W)


def intra_so3conv(feats, W, intra_idx):
    """IntraSO3Conv.forward: vgtk/vgtk/so3conv/modules.py:197-200"""
    return basic_conv(intra_group(intra_idx, feats), W)


# ---------------------------------------------------------------------- blocks
def _norm(x, kind, weight=None, bias=None, eps=1e-5):
    """BatchNorm2d in training mode (batch statistics) or InstanceNorm2d(affine=False);
    SPConvNets/utils/base_so3conv.py:43,107,193."""
    if kind == "BatchNorm2d":
        return F.batch_norm(x, None, None, weight, bias, True, 0.1, eps)
    return F.instance_norm(x, eps=eps)


def separable_block(xyz, feats, prm, args, intra_idx, anchors, kernels):
    """SeparableSO3ConvBlock.forward (SPConvNets/utils/base_so3conv.py:197-212) with the
    InterSO3ConvBlock (:116-126) and IntraSO3ConvBlock (:52-62) it calls; training mode, dropout 0.
    prm: dict of tensors {inter_W, inter_bn_w, inter_bn_b, intra_W, skip_w, skip_b, bn_w, bn_b}."""
    norm = args.get("norm")
    skip = feats
    _, _, sample_idx, new_xyz, x = inter_so3conv(xyz, feats, prm["inter_W"], anchors, kernels, args["stride"],
                                                 args["n_neighbor"], args["radius"], args["sigma"],
                                                 args["lazy_sample"])
    x = F.leaky_relu(_norm(x, norm, prm.get("inter_bn_w"), prm.get("inter_bn_b")))
    if intra_idx is not None:
        x = intra_so3conv(x, prm["intra_W"], intra_idx)
        x = F.leaky_relu(_norm(x, None))
    if args["stride"] > 1:
        b, c, _, a = skip.shape
        index = sample_idx.long().view(b, 1, -1, 1).expand(b, c, sample_idx.shape[1], a)
        skip = torch.gather(skip, 2, index)
    skip = F.conv2d(skip, prm["skip_w"], prm["skip_b"])
    skip = F.leaky_relu(_norm(skip, norm, prm.get("bn_w"), prm.get("bn_b")))
    return new_xyz, x + skip


def backbone_forward(x, layers, dtype=torch.float32):
    """ClsSO3ConvModel.forward minus the head (SPConvNets/models/cls_so3net_pn.py:27-33;
    preprocess_input base_so3conv.py:16-23 with add_center=False).
    x [b,n,3]; layers: list of (prm, args, intra_idx, anchors, kernels)."""
    b, n, _ = x.shape
    xyz = x.permute(0, 2, 1).contiguous()
    na = layers[0][3].shape[0]
    feats = torch.ones(b, 1, n, na, dtype=dtype, device=x.device)  # dtype=float64 gives the "true" value the fp32 paths round
    for prm, args, intra_idx, anchors, kernels in layers:
        xyz, feats = separable_block(xyz, feats, prm, args, intra_idx, anchors, kernels)
    return xyz, feats


def cls_head(xyz, feats, hp):
    """ClsOutBlockPointnet.forward with 'max' pooling, training mode (SPConvNets/utils/base_so3conv.py:404-448)
    incl. PointnetSO3Conv (vgtk/vgtk/so3conv/modules.py:218-235).  hp: dict of tensors
    {lin_w[i], lin_b[i], bn_w[i], bn_b[i] (lists), pn_w, pn_b, bn1_w, bn1_b, fc_w, fc_b, anchors}."""
    x = feats
    for w, b, gw, gb in zip(hp["lin_w"], hp["lin_b"], hp["bn_w"], hp["bn_b"]):
        x = F.relu(F.batch_norm(F.conv2d(x, w, b), None, None, gw, gb, True, 0.1, 1e-5))
    out_feat = x
    c = xyz - xyz.mean(2, keepdim=True)
    xyzr = torch.einsum("aji,bjn->bina", hp["anchors"].to(x.dtype), c.to(x.dtype))
    x = F.conv2d(torch.cat([x, xyzr], 1), hp["pn_w"], hp["pn_b"]).max(2)[0]
    x = F.relu(F.batch_norm(x, None, None, hp["bn1_w"], hp["bn1_b"], True, 0.1, 1e-5))
    return F.linear(x.max(2)[0], hp["fc_w"], hp["fc_b"]), out_feat


def head_from_module(outblock):
    """Tensors of an `epn_pointcloud_b200.heads.ClsOutBlockPointnet` (or the reference's) for cls_head()."""
    sd = {k: v.detach().cpu().float() for k, v in outblock.state_dict().items()}
    n = len(outblock.linear)
    return {"lin_w": [sd["linear.%d.weight" % i] for i in range(n)], "lin_b": [sd["linear.%d.bias" % i] for i in range(n)],
            "bn_w": [sd["norm.%d.weight" % i] for i in range(n)], "bn_b": [sd["norm.%d.bias" % i] for i in range(n)],
            "pn_w": sd["pointnet.embed.weight"], "pn_b": sd["pointnet.embed.bias"], "anchors": sd["pointnet.anchors"],
            "bn1_w": sd["norm.%d.weight" % n], "bn1_b": sd["norm.%d.bias" % n], "fc_w": sd["fc2.weight"],
            "fc_b": sd["fc2.bias"]}


def layers_from_module(backbone):
    """Pull (prm, args, ...) out of an `epn_pointcloud_b200.blocks.SO3ConvBackbone` (or the
    reference's ModuleList of BasicSO3ConvBlock) living on any device -> CPU tensors."""
    layers = []
    for blk in backbone.backbone:
        for conv, param in zip(blk.blocks, blk.params):
            assert param["type"] == "separable_block"
            sd = {k: v.detach().cpu().float() for k, v in conv.state_dict().items()}
            prm = {"inter_W": sd["inter_conv.conv.basic_conv.W"], "skip_w": sd["skip_conv.weight"],
                   "skip_b": sd["skip_conv.bias"]}
            if "inter_conv.norm.weight" in sd:
                prm["inter_bn_w"], prm["inter_bn_b"] = sd["inter_conv.norm.weight"], sd["inter_conv.norm.bias"]
            if "norm.weight" in sd:
                prm["bn_w"], prm["bn_b"] = sd["norm.weight"], sd["norm.bias"]
            intra_idx = None
            if "intra_conv.conv.basic_conv.W" in sd:
                prm["intra_W"] = sd["intra_conv.conv.basic_conv.W"]
                intra_idx = conv.state_dict()["intra_conv.conv.intra_idx"].cpu()
            layers.append((prm, param["args"], intra_idx, sd["inter_conv.conv.anchors"],
                           sd["inter_conv.conv.kernels"]))
    return layers


# ------------------------------------------------------- heads of the other two shipped models (8 f2)
def _pointnet_so3(xyz, feats, w, b, anchors):
    """PointnetSO3Conv.forward (vgtk/vgtk/so3conv/modules.py:218-235): centred coordinates rotated into every anchor
    frame, concatenated to the features, 1x1 conv, max over points -> [nb, c_out, na]."""
    c = xyz - xyz.mean(2, keepdim=True)
    na = feats.shape[3]
    if na == 1:
        x = torch.cat([feats, c[..., None]], 1)
    else:
        x = torch.cat([feats, torch.einsum("aji,bjn->bina", anchors.to(feats.dtype), c.to(feats.dtype))], 1)
    return F.conv2d(x, w, b).max(2)[0]


def inv_head(xyz, feats, hp):
    """InvOutBlockMVD.forward (SPConvNets/utils/base_so3conv.py:596-613).  hp: {att0_w, att0_b, att2_w, att2_b,
    pn_w, pn_b, anchors} -> (descriptor [nb, c_out] L2-normalised, attention [nb, c, np, na])."""
    nb = feats.shape[0]
    attn = F.conv2d(F.relu(F.conv2d(feats, hp["att0_w"], hp["att0_b"])), hp["att2_w"], hp["att2_b"])
    attn = F.softmax(attn, dim=3)
    x = (feats * attn).sum(-1, keepdim=True)
    x = _pointnet_so3(xyz, x, hp["pn_w"], hp["pn_b"], hp["anchors"]).view(nb, -1)
    return F.normalize(x, p=2, dim=1), attn


def rel_head(f1, f2, x1, x2, hp):
    """RelSO3OutBlockR.forward (SPConvNets/utils/base_so3conv.py:696-730).  hp: {pn_w, pn_b, anchors, lin_w[], lin_b[],
    att_w, att_b, reg_w, reg_b, temperature} -> (confidence [nb, na, na], y [nb, n_out, na, na])."""
    p1 = F.relu(_pointnet_so3(x1, f1, hp["pn_w"], hp["pn_b"], hp["anchors"]))
    p2 = F.relu(_pointnet_so3(x2, f2, hp["pn_w"], hp["pn_b"], hp["anchors"]))
    nb, na = p1.shape[0], p1.shape[2]
    x = torch.cat((p1.unsqueeze(-2).expand(-1, -1, na, -1), p2.unsqueeze(-1).expand(-1, -1, -1, na)), 1).contiguous()
    for w, b in zip(hp["lin_w"], hp["lin_b"]):
        x = F.relu(F.conv2d(x, w, b))
    conf = F.softmax(F.conv2d(x, hp["att_w"], hp["att_b"]).view(nb, na, na) * hp["temperature"], dim=1)
    return conf, F.conv2d(x, hp["reg_w"], hp["reg_b"])


def inv_head_from_state(sd):
    """state_dict of an InvOutBlockMVD (reference or mirror; keys relative to the block)."""
    return {"att0_w": sd["attention_layer.0.weight"], "att0_b": sd["attention_layer.0.bias"],
            "att2_w": sd["attention_layer.2.weight"], "att2_b": sd["attention_layer.2.bias"],
            "pn_w": sd["pointnet.embed.weight"], "pn_b": sd["pointnet.embed.bias"], "anchors": sd["pointnet.anchors"]}


def rel_head_from_state(sd, temperature):
    n = len([k for k in sd if k.startswith("linear.") and k.endswith(".weight")])
    return {"pn_w": sd["pointnet.embed.weight"], "pn_b": sd["pointnet.embed.bias"], "anchors": sd["pointnet.anchors"],
            "lin_w": [sd["linear.%d.weight" % i] for i in range(n)], "lin_b": [sd["linear.%d.bias" % i] for i in range(n)],
            "att_w": sd["attention_layer.weight"], "att_b": sd["attention_layer.bias"],
            "reg_w": sd["regressor_layer.weight"], "reg_b": sd["regressor_layer.bias"], "temperature": temperature}
