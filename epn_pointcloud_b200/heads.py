"""Output head of the classification model and the full model wrapper.

Mirrors, with identical constructor arguments / sub-module names / state_dict keys:
  vgtk/vgtk/so3conv/modules.py:203-235            PointnetSO3Conv
  SPConvNets/utils/base_so3conv.py:358-448        ClsOutBlockPointnet
  SPConvNets/models/cls_so3net_pn.py:16-39,41-167 ClsSO3ConvModel / build_model

The head works on tiny tensors ([B, C, 64, 60] and smaller), so it is plain PyTorch plumbing around
one library op: every 1x1 convolution runs through the library's fp32-faithful channel GEMM
(`BasicSO3Conv` with kernel size 1) instead of cuDNN, which would silently use TF32.
"""
import torch
import torch.nn as nn
import torch.nn.functional as F

from . import functional as L
from . import modules as sptk
from .blocks import BasicSO3ConvBlock, backbone_params, cls_backbone_params, fwd_operands, norm_act, preprocess_input


def conv1x1(conv, x, add_bias=True):
    """nn.Conv2d(c_in, c_out, 1) applied to x [b, c_in, p, a] through the library GEMM (parameters stay in
    the nn.Conv2d so checkpoints keep their keys).  add_bias=False leaves the bias to a following norm_act."""
    w = conv.weight.view(conv.out_channels, conv.in_channels)
    with fwd_operands(x):
        y = sptk._BasicConvFn.apply(x.contiguous().unsqueeze(2), w)
    return y if (conv.bias is None or not add_bias) else y + conv.bias.view(1, -1, 1, 1)


class PointnetSO3Conv(nn.Module):
    """Equivariant PointNet aggregation: append the anchor-rotated, centred coordinates to the features,
    1x1 conv, max over points -> [nb, dim_out, na]  (so3conv/modules.py:203-235)."""

    def __init__(self, dim_in, dim_out, kanchor=60):
        super().__init__()
        anchors = L.get_anchors(kanchor)
        self.dim_in = dim_in + 3
        self.dim_out = dim_out
        self.embed = nn.Conv2d(self.dim_in, self.dim_out, 1)
        self.register_buffer("anchors", torch.from_numpy(anchors))

    def forward(self, x):
        xyz = x.xyz - x.xyz.mean(2, keepdim=True)
        na = x.feats.shape[3]
        if na == 1:
            feats = torch.cat([x.feats, xyz[..., None]], 1)
        else:
            xyzr = torch.einsum("aji,bjn->bina", self.anchors, xyz)
            feats = torch.cat([x.feats, xyzr], 1)
        feats = conv1x1(self.embed, feats)
        return torch.max(feats, 2)[0]


class ClsOutBlockPointnet(nn.Module):
    """1x1 conv + BN + relu stack, PointnetSO3Conv, BN1d + relu, anchor pooling, linear classifier
    (base_so3conv.py:358-448).  Returns (logits [nb, k], per-anchor features)."""

    def __init__(self, params, norm=None, debug=False):
        super().__init__()
        c_in = params["dim_in"]
        self.outDim = params["k"]
        na = params["kanchor"]
        self.linear = nn.ModuleList()
        self.norm = nn.ModuleList()
        for c in params["mlp"]:
            self.linear.append(nn.Conv2d(c_in, c, 1))
            self.norm.append(nn.BatchNorm2d(c))
            c_in = c
        self.pooling_method = params.get("pooling", "max")
        if self.pooling_method == "attention":
            self.temperature = params["temperature"]
            self.attention_layer = nn.Conv1d(c_in, 1, 1)
        self.pointnet = PointnetSO3Conv(c_in, c_in, na)
        self.norm.append(nn.BatchNorm1d(c_in))
        self.fc2 = nn.Linear(c_in, self.outDim)
        self.debug = debug

    def forward(self, x, label=None):
        x_out = x.feats
        if self.debug:
            return x_out[:, :40].mean(-1).mean(-1), None
        norm_cnt = 0
        for linear in self.linear:
            # conv bias + BatchNorm2d + relu as one fused pass (blocks.norm_act; the bias cancels in the batch statistics)
            x_out = norm_act(self.norm[norm_cnt], conv1x1(linear, x_out, add_bias=False), F.relu, bias=linear.bias)
            norm_cnt += 1
        out_feat = x_out
        x_out = self.pointnet(sptk.SphericalPointCloud(x.xyz, out_feat, x.anchors))
        x_out = F.relu(self.norm[norm_cnt](x_out))
        if self.pooling_method == "mean":
            x_out = x_out.mean(dim=2)
        elif self.pooling_method == "debug":
            x_out = x_out[..., 0].mean(2)
        elif self.pooling_method == "max":
            x_out = x_out.max(2)[0]
        elif self.pooling_method.startswith("attention"):
            out_feat = self.attention_layer(x_out)
            confidence = F.softmax(out_feat * self.temperature, dim=2)
            x_out = (x_out * confidence).sum(-1)
        else:
            raise NotImplementedError("Pooling mode %s is not implemented!" % self.pooling_method)
        return self.fc2(x_out), out_feat.squeeze()


class ClsSO3ConvModel(nn.Module):
    """ModelNet40 classification network: backbone of BasicSO3ConvBlocks + ClsOutBlockPointnet
    (cls_so3net_pn.py:16-39)."""

    def __init__(self, params):
        super().__init__()
        self.backbone = nn.ModuleList([BasicSO3ConvBlock(bp) for bp in params["backbone"]])
        self.outblock = ClsOutBlockPointnet(params["outblock"])
        self.na_in = params["na"]
        self.invariance = True

    def forward(self, x, rlabel=None):
        x = preprocess_input(x, self.na_in, False)
        for block in self.backbone:
            x = block(x)
        return self.outblock(x, rlabel)

    def get_anchor(self):
        return self.backbone[-1].get_anchor()


def cls_model_params(input_num=1024, kanchor=60, dropout_rate=0.0, so3_pooling="max", temperature=3.0,
                     out_mlps=(256,), **backbone_kwargs):
    """The `params` dict of cls_so3net_pn.build_model (cls_so3net_pn.py:41-160)."""
    backbone = cls_backbone_params(input_num, kanchor, dropout_rate, **backbone_kwargs)
    dim_in = backbone[-1][-1]["args"]["dim_out"]
    return {"name": "Invariant ZPConv Model", "backbone": backbone, "na": kanchor,
            "outblock": {"dim_in": dim_in, "mlp": list(out_mlps), "fc": [64], "k": 40, "pooling": so3_pooling,
                         "temperature": temperature, "kanchor": kanchor}}


# ------------------------------------------------------------------ 3DMatch descriptor model
class InvOutBlockMVD(nn.Module):
    """Attention over the anchors, then equivariant PointNet pooling to one rotation-invariant descriptor
    (SPConvNets/utils/base_so3conv.py:572-613)."""

    def __init__(self, params, norm=None):
        super().__init__()
        c_in = params["dim_in"]
        c_out = params["mlp"][-1]
        na = params["kanchor"]
        self.temperature = params["temperature"]
        self.attention_layer = nn.Sequential(nn.Conv2d(c_in, c_in, 1), nn.ReLU(inplace=True), nn.Conv2d(c_in, c_in, 1))
        self.pooling_method = params.get("pooling", "max")
        self.pointnet = PointnetSO3Conv(c_in, c_out, na)

    def forward(self, x):
        nb = x.feats.shape[0]
        attn = conv1x1(self.attention_layer[2], F.relu(conv1x1(self.attention_layer[0], x.feats)))
        attn = F.softmax(attn, dim=3)
        x_out = (x.feats * attn).sum(-1, keepdim=True)
        x_out = self.pointnet(sptk.SphericalPointCloud(x.xyz, x_out, None)).view(nb, -1)
        return F.normalize(x_out, p=2, dim=1), attn


class InvSO3ConvModel(nn.Module):
    """3DMatch local-patch descriptor network (SPConvNets/models/inv_so3net_pn.py:15-41)."""

    def __init__(self, params):
        super().__init__()
        self.backbone = nn.ModuleList([BasicSO3ConvBlock(bp) for bp in params["backbone"]])
        self.outblock = InvOutBlockMVD(params["outblock"])
        self.na_in = params["na"]
        self.invariance = True

    def forward(self, x):
        x = preprocess_input(x, self.na_in, False)
        for block in self.backbone:
            x = block(x)
        return self.outblock(x)

    def get_anchor(self):
        return self.backbone[-1].get_anchor()


def inv_model_params(input_num=2048, kanchor=60, dropout_rate=0.0, so3_pooling="max", temperature=3.0,
                     search_radius=0.4, mlps=((32, 32), (64, 64), (128, 128), (128, 128)), out_mlps=(128, 64),
                     strides=(2, 2, 2, 2), **kw):
    """The `params` dict of inv_so3net_pn.build_model (inv_so3net_pn.py:43-163)."""
    backbone = backbone_params(input_num, kanchor, dropout_rate, mlps=mlps, strides=strides, sampling_ratio=0.8,
                               input_radius=search_radius, norm=None, sigma_rule="stride", scale_first_neighbor=True, **kw)
    dim_in = backbone[-1][-1]["args"]["dim_out"]
    return {"name": "Invariant ZPConv Model", "backbone": backbone, "na": kanchor,
            "outblock": {"dim_in": dim_in, "mlp": list(out_mlps), "pooling": so3_pooling, "temperature": temperature,
                         "kanchor": kanchor}}


# ------------------------------------------------------------------ relative rotation model
class RelSO3OutBlockR(nn.Module):
    """Pairwise anchor-alignment head: pooled features of the two clouds are combined into a [2C, 60, 60] tensor,
    1x1 convs, a confidence over the source anchors and one rotation residual per anchor pair
    (SPConvNets/utils/base_so3conv.py:661-730)."""

    def __init__(self, params, norm=None):
        super().__init__()
        c_in = params["dim_in"]
        mlp = params["mlp"]
        na = params["kanchor"]
        self.pointnet = PointnetSO3Conv(c_in, c_in, na)
        c_in = c_in * 2
        self.linear = nn.ModuleList()
        self.temperature = params["temperature"]
        rp = params["representation"]
        if rp == "quat":
            self.out_channel = 4
        elif rp == "ortho6d":
            self.out_channel = 6
        else:
            raise KeyError("Unrecognized representation of rotation: %s" % rp)
        self.attention_layer = nn.Conv2d(mlp[-1], 1, (1, 1))
        self.regressor_layer = nn.Conv2d(mlp[-1], self.out_channel, (1, 1))
        for c in mlp:
            self.linear.append(nn.Conv2d(c_in, c, (1, 1)))
            c_in = c

    def forward(self, f1, f2, x1, x2):
        f1 = self._pooling(sptk.SphericalPointCloud(x1, f1, None))
        f2 = self._pooling(sptk.SphericalPointCloud(x2, f2, None))
        nb, na = f1.shape[0], f1.shape[2]
        f2_expand = f2.unsqueeze(-1).expand(-1, -1, -1, na).contiguous()
        f1_expand = f1.unsqueeze(-2).expand(-1, -1, na, -1).contiguous()
        x_out = torch.cat((f1_expand, f2_expand), 1)
        for linear in self.linear:
            x_out = F.relu(conv1x1(linear, x_out))
        attention_wts = conv1x1(self.attention_layer, x_out).view(nb, na, na)
        confidence = F.softmax(attention_wts * self.temperature, dim=1)
        y = conv1x1(self.regressor_layer, x_out)
        return confidence, y

    def _pooling(self, x):
        return F.relu(self.pointnet(x))


class RegSO3ConvModel(nn.Module):
    """ModelNet40 relative rotation estimation network; input [nb, 2, np, 3] = (source, target) pairs
    (SPConvNets/models/reg_so3net.py:16-52)."""

    def __init__(self, params):
        super().__init__()
        self.backbone = nn.ModuleList([BasicSO3ConvBlock(bp) for bp in params["backbone"]])
        self.outblock = RelSO3OutBlockR(params["outblock"])
        self.na_in = params["na"]
        self.invariance = True

    def forward(self, x):
        x = torch.cat((x[:, 0], x[:, 1]), dim=0)
        x = preprocess_input(x, self.na_in, False)
        for block in self.backbone:
            x = block(x)
        f1, f2 = torch.chunk(x.feats, 2, dim=0)
        x1, x2 = torch.chunk(x.xyz, 2, dim=0)
        return self.outblock(f1, f2, x1, x2)

    def get_anchor(self):
        return self.backbone[-1].get_anchor()


def reg_model_params(input_num=1024, kanchor=60, dropout_rate=0.0, temperature=3.0, representation="quat",
                     mlps=((32, 32), (64, 64), (128, 128), (256,)), out_mlps=(256, 128, 64), strides=(2, 2, 2, 2), **kw):
    """The `params` dict of reg_so3net.build_model (reg_so3net.py:54-171)."""
    backbone = backbone_params(input_num, kanchor, dropout_rate, mlps=mlps, strides=strides, sampling_ratio=0.8,
                               norm=None, **kw)
    dim_in = backbone[-1][-1]["args"]["dim_out"]
    return {"name": "Invariant ZPConv Model", "backbone": backbone, "na": kanchor,
            "outblock": {"dim_in": dim_in, "mlp": list(out_mlps), "fc": [64], "k": 40, "kanchor": kanchor,
                         "representation": representation, "temperature": temperature}}
