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
from .blocks import BasicSO3ConvBlock, cls_backbone_params, preprocess_input


def conv1x1(conv, x):
    """nn.Conv2d(c_in, c_out, 1) applied to x [b, c_in, p, a] through the library GEMM (parameters stay in
    the nn.Conv2d so checkpoints keep their keys)."""
    w = conv.weight.view(conv.out_channels, conv.in_channels)
    y = sptk._BasicConvFn.apply(x.contiguous().unsqueeze(2), w)
    return y if conv.bias is None else y + conv.bias.view(1, -1, 1, 1)


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
            x_out = F.relu(self.norm[norm_cnt](conv1x1(linear, x_out)))
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
