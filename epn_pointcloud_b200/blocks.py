"""Block wrappers around the fused convs -- mirror of SPConvNets/utils/base_so3conv.py:16-215
(`preprocess_input`, `InterSO3ConvBlock`, `IntraSO3ConvBlock`, `SeparableSO3ConvBlock`,
`BasicSO3ConvBlock`) with identical constructor arguments and sub-module names, so a
reference `state_dict` loads key for key, plus the backbone parameter arithmetic of the three
shipped models (SPConvNets/models/cls_so3net_pn.py:41-150, reg_so3net.py:76-171,
inv_so3net_pn.py:66-163).
"""
import torch
import torch.nn as nn
import torch.nn.functional as F

from . import functional as L
from . import modules as sptk
from . import ops


class _NormActFn(torch.autograd.Function):
    """leaky_relu(norm(x)) [+ residual] as one library op (epn_norm_act_*); saves x and the (mean, rstd) statistics."""

    @staticmethod
    def forward(ctx, x, gamma, beta, mode, eps, slope, residual=None, cancelled_bias=None):
        # cancelled_bias: a per-channel constant added to x upstream (the skip convolution's bias).  The normalisation
        # removes it again, so it is never added; it only takes part in autograd, where its exact gradient is zero.
        ctx.bias_shape = None if cancelled_bias is None else cancelled_bias.shape
        x = x.contiguous()
        if residual is not None:
            residual = residual.contiguous()
        y, stats = ops.norm_act_fwd(x, gamma, beta, mode, eps, slope, residual)
        ctx.save_for_backward(x, stats, gamma, beta)
        ctx.mode, ctx.slope = mode, slope
        ctx.mark_non_differentiable(stats)
        return y, stats

    @staticmethod
    def backward(ctx, dy, _dstats):
        x, stats, gamma, beta = ctx.saved_tensors
        dy = dy.contiguous()
        dx, dgamma, dbeta = ops.norm_act_bwd(dy, x, gamma, beta, stats, ctx.mode, ctx.slope)
        dbias = dy.new_zeros(ctx.bias_shape) if (ctx.bias_shape is not None and ctx.needs_input_grad[7]) else None
        return dx, dgamma, dbeta, None, None, None, (dy if ctx.needs_input_grad[6] else None), dbias


# ------------------------------------------------------------------ operand format of inference forwards
# Under torch.no_grad() the convs of these wrappers run their tensor-core GEMMs on fp16 hi/lo operands (22 significand
# bits instead of bf16 hi/lo's 16, same speed; include/epn_b200.h, epn_set_forward_operands) -- but ONLY when the
# conv's input is known to be a normalised activation: a tensor produced by norm_act below (tagged `_epn_unit`) or the
# constant occupancy features.  Anything of unknown provenance keeps the range-safe bf16 operands.
_INFERENCE_OPERANDS = "f16"


def set_inference_operands(fmt):
    """'f16' (default) or 'bf16': operand format of no_grad forwards on normalised activations."""
    global _INFERENCE_OPERANDS
    if fmt not in ("f16", "bf16"):
        raise ValueError(fmt)
    _INFERENCE_OPERANDS = fmt


def _mark_unit(t):
    t._epn_unit = True
    return t


def fwd_operands(feats=None, occupancy=False):
    """Context manager for ONE conv forward whose input features are `feats`."""
    ok = (not torch.is_grad_enabled() and _INFERENCE_OPERANDS == "f16" and
          (occupancy or (feats is not None and getattr(feats, "_epn_unit", False))))
    return ops.forward_operands("f16" if ok else "bf16")


def _bn_track(norm, stats, count, bias=None):
    """The running-statistics bookkeeping of nn.BatchNorm2d.forward in training mode, from the batch (mean, rstd) the
    fused kernels computed; `bias`: a per-channel constant that was left out of the kernel's input (see norm_act)."""
    if not norm.track_running_stats:
        return
    with torch.no_grad():
        if (stats.is_cuda and norm.running_mean.dtype == torch.float32 and norm.running_mean.is_contiguous()
                and norm.running_var.is_contiguous() and norm.num_batches_tracked.dtype == torch.int64):
            ops.bn_track(stats, None if bias is None else bias.detach().contiguous(), norm.running_mean, norm.running_var,
                         norm.num_batches_tracked, count, norm.momentum, norm.eps)   # one launch, graph-capturable
            return
        norm.num_batches_tracked += 1
        mean = stats[0] if bias is None else stats[0] + bias.detach()
        var = (1.0 / (stats[1] * stats[1]) - norm.eps) * (count / max(count - 1, 1))
        if norm.momentum is not None:
            m = norm.momentum
            norm.running_mean.mul_(1 - m).add_(mean, alpha=m)
            norm.running_var.mul_(1 - m).add_(var, alpha=m)
        else:  # cumulative moving average: the factor stays on the device (no host sync, graph-capturable)
            m = 1.0 / norm.num_batches_tracked.to(torch.float32)
            norm.running_mean.add_((mean - norm.running_mean) * m)
            norm.running_var.add_((var - norm.running_var) * m)


class _NormIntraFn(torch.autograd.Function):
    """IntraSO3Conv(leaky_relu(norm(x))) with the normalised activation never written to memory: the statistics come
    from one pass over x, the normalisation + activation are applied while the intra conv builds its operand tiles
    (ops.intra_so3conv_fwd_norm; SURVEY 8 row f1, base_so3conv.py:116-126 -> :52-62).  Backward: intra conv backward
    from the kept tiles (dW) and the permuted GEMM (gradient of the activation), then the norm backward from x and
    the statistics -- neither needs the activation.  Falls back to the two separate ops (keeping the activation for
    the backward's re-gather) when the shape is not covered or no memory is granted for the kept tiles."""

    @staticmethod
    def forward(ctx, x, gamma, beta, W, intra_idx, mode, eps, slope, given_stats, training):
        x, W = x.contiguous(), W.contiguous()
        kmode = 1 if mode == 2 else mode
        stats = given_stats if mode == 2 else ops.norm_stats(x, mode, eps)
        keep = bool(training) and bool(ctx.needs_input_grad[0] or ctx.needs_input_grad[3])
        res = ops.intra_so3conv_fwd_norm(x, stats, gamma, beta, kmode, slope, intra_idx, W, keep_grouped=keep)
        y = None
        if res is None:
            y = ops.norm_act_fwd(x, gamma, beta, mode, eps, slope, stats=stats if mode == 2 else None)[0]
            res = ops.intra_so3conv_fwd(y, intra_idx, W, keep_grouped=keep)
        out, ctx.grouped = res if keep else (res, None)
        ctx.save_for_backward(x, stats, gamma, beta, W, intra_idx, *([y] if (y is not None and ctx.grouped is None) else []))
        ctx.mode, ctx.slope = kmode, slope
        ctx.mark_non_differentiable(stats)
        return out, stats

    @staticmethod
    def backward(ctx, dout, _dstats):
        saved = list(ctx.saved_tensors)
        x, stats, gamma, beta, W, intra_idx = saved[:6]
        y = saved[6] if len(saved) > 6 else x       # only read when there are no kept tiles
        dy, dW = ops.intra_so3conv_bwd(dout.contiguous(), y, intra_idx, W, True, ctx.needs_input_grad[3], grouped=ctx.grouped)
        ctx.grouped = None
        dx, dgamma, dbeta = ops.norm_act_bwd(dy, x, gamma, beta, stats, ctx.mode, ctx.slope)
        return dx, dgamma, dbeta, dW, None, None, None, None, None, None


def norm_intra(norm, x, act, intra):
    """`intra(act(norm(x)))` for x = the raw output of a conv and `intra` an IntraSO3Conv module, as one fused op when
    (norm, act) is one of the combinations norm_act fuses; returns None when it is not (caller runs the two ops)."""
    if not (x.is_cuda and x.dtype == torch.float32 and act is F.leaky_relu and x.dim() == 4 and _FUSE_NORM_INTRA):
        return None
    if torch.is_grad_enabled() and not _FUSE_NORM_INTRA_TRAINING:
        return None
    W, idx = intra.basic_conv.W, intra._intra_idx32
    if isinstance(norm, nn.InstanceNorm2d) and not norm.affine and not norm.track_running_stats:
        return _NormIntraFn.apply(x, None, None, W, idx, 0, norm.eps, 0.01, None, torch.is_grad_enabled())[0]
    if isinstance(norm, nn.BatchNorm2d) and norm.training and norm.affine:
        out, stats = _NormIntraFn.apply(x, norm.weight, norm.bias, W, idx, 1, norm.eps, 0.01, None, torch.is_grad_enabled())
        _bn_track(norm, stats, x.numel() // x.shape[1])
        return out
    if (isinstance(norm, nn.BatchNorm2d) and not norm.training and norm.track_running_stats and norm.running_mean is not None
            and not torch.is_grad_enabled()):
        stats = torch.stack((norm.running_mean, torch.rsqrt(norm.running_var + norm.eps))).contiguous()
        return _NormIntraFn.apply(x, norm.weight, norm.bias, W, idx, 2, norm.eps, 0.01, stats, False)[0]
    return None


_FUSE_NORM_INTRA = True   # tests switch it off to compare the fused pair with the two separate ops
# Under autograd the fused pair is correct (tested) but NOT faster on the B200: the training forward builds the kept
# operand tiles with a 12x gather (intra_tiles_rows_kernel, HBM-write-bound), and normalising every gathered element
# there doubles that kernel's instruction count (measured 3.95 -> 8.0 ms per step against 0.5 ms saved in the norm
# kernels).  So by default only no_grad forwards fuse (29.35 -> 28.9 ms inference forward).
_FUSE_NORM_INTRA_TRAINING = False


def norm_act(norm, x, act, residual=None, bias=None):
    return _mark_unit(_norm_act(norm, x, act, residual, bias))


def _norm_act(norm, x, act, residual=None, bias=None):
    """`act(norm(x + bias)) + residual` of the block wrappers (base_so3conv.py:55-57,119-125,209-211).  The two
    combinations every shipped model uses on CUDA -- InstanceNorm2d(affine=False) / training-mode BatchNorm2d followed
    by leaky_relu -- run as one fused library op; anything else goes through the torch modules.

    `bias` [c] is the bias of the 1x1 skip convolution that produced x.  A normalisation with statistics of x itself
    subtracts the per-channel mean, so a per-channel constant cancels exactly: the fused paths never add it (one
    full pass over the tensor saved; its gradient is identically zero) and only fold it into BatchNorm's running
    mean, which does see it."""
    fusable = x.is_cuda and x.dtype == torch.float32 and (act is F.leaky_relu or act is F.relu) and x.dim() == 4
    slope = 0.01 if act is F.leaky_relu else 0.0    # F.leaky_relu's default slope; relu = slope 0
    if fusable and isinstance(norm, nn.InstanceNorm2d) and not norm.affine and not norm.track_running_stats:
        return _NormActFn.apply(x, None, None, 0, norm.eps, slope, residual, bias)[0]
    if fusable and isinstance(norm, nn.BatchNorm2d) and norm.training and norm.affine:
        y, stats = _NormActFn.apply(x, norm.weight, norm.bias, 1, norm.eps, slope, residual, bias)
        _bn_track(norm, stats, x.numel() // x.shape[1], bias)
        return y
    if (fusable and isinstance(norm, nn.BatchNorm2d) and not norm.training and norm.track_running_stats
            and norm.running_mean is not None and not (torch.is_grad_enabled() and x.requires_grad)):
        # evaluation mode: per-channel affine map from the running statistics, one pass (no statistics kernels)
        mean = norm.running_mean if bias is None else norm.running_mean - bias.detach()
        stats = torch.stack((mean, torch.rsqrt(norm.running_var + norm.eps))).contiguous()
        return ops.norm_act_fwd(x, norm.weight, norm.bias, 2, norm.eps, slope, residual, stats=stats)[0]
    if bias is not None:
        x = x + bias.view(1, -1, 1, 1)
    out = norm(x)
    out = act(out) if act is not None else out
    return out if residual is None else out + residual


def preprocess_input(x, na, add_center=True):
    """[nb, np, 3] -> SphericalPointCloud(xyz [nb,3,np], occupancy feats [nb,1,np,na])
    (base_so3conv.py:16-23)."""
    if x.shape[2] != 3:
        raise NotImplementedError("normals input: broken upstream (so3conv/functional.py:35-36)")
    if add_center:
        center = x.mean(1, keepdim=True)
        x = torch.cat((center, x), dim=1)[:, :-1]
    xyz = x[:, :, :3].permute(0, 2, 1).contiguous()
    if add_center:  # feature of the dummy centre point is zeroed: needs the real tensor
        return sptk.SphericalPointCloud(xyz, L.get_occupancy_features(x, na, True), None)
    return sptk.SphericalPointCloud(xyz, None, None, occupancy=(x.shape[0], x.shape[1], na))


class IntraSO3ConvBlock(nn.Module):
    """base_so3conv.py:32-62"""

    def __init__(self, dim_in, dim_out, norm=None, activation="relu", dropout_rate=0):
        super().__init__()
        if norm is not None:
            norm = getattr(nn, norm)
        self.conv = sptk.IntraSO3Conv(dim_in, dim_out)
        self.norm = nn.InstanceNorm2d(dim_out, affine=False) if norm is None else norm(dim_out)
        self.relu = None if activation is None else getattr(F, activation)
        self.dropout = nn.Dropout(dropout_rate) if dropout_rate > 0 else None

    def forward(self, x):
        with fwd_operands(x.feats):
            x = self.conv(x)
        feat = norm_act(self.norm, x.feats, self.relu)
        if self.training and self.dropout is not None:
            feat = self.dropout(feat)
        return sptk.SphericalPointCloud(x.xyz, feat, x.anchors)


class InterSO3ConvBlock(nn.Module):
    """base_so3conv.py:88-126"""

    def __init__(self, dim_in, dim_out, kernel_size, stride, radius, sigma, n_neighbor, multiplier, kanchor=60,
                 lazy_sample=None, norm=None, activation="relu", pooling="none", dropout_rate=0):
        super().__init__()
        if lazy_sample is None:
            lazy_sample = True
        if norm is not None:
            norm = getattr(nn, norm)
        pooling_method = None if pooling == "none" else pooling
        self.conv = sptk.InterSO3Conv(dim_in, dim_out, kernel_size, stride, radius, sigma, n_neighbor,
                                      kanchor=kanchor, lazy_sample=lazy_sample, pooling=pooling_method)
        self.norm = nn.InstanceNorm2d(dim_out, affine=False) if norm is None else norm(dim_out)
        self.relu = None if activation is None else getattr(F, activation)
        self.dropout = nn.Dropout(dropout_rate) if dropout_rate > 0 else None

    def forward(self, x, inter_idx=None, inter_w=None):
        occ = bool(getattr(x, "is_occupancy", False))
        with fwd_operands(None if occ else x.feats, occupancy=occ):
            inter_idx, inter_w, sample_idx, x = self.conv(x, inter_idx, inter_w)
        feat = norm_act(self.norm, x.feats, self.relu)
        if self.training and self.dropout is not None:
            feat = self.dropout(feat)
        return inter_idx, inter_w, sample_idx, sptk.SphericalPointCloud(x.xyz, feat, x.anchors)


class SeparableSO3ConvBlock(nn.Module):
    """inter conv -> intra conv, plus the 1x1-conv skip branch (base_so3conv.py:168-215)"""

    def __init__(self, params):
        super().__init__()
        dim_in = params["dim_in"]
        dim_out = params["dim_out"]
        norm = getattr(nn, params["norm"]) if "norm" in params.keys() else None
        self.use_intra = params["kanchor"] > 1
        self.inter_conv = InterSO3ConvBlock(**params)
        intra_args = {"dim_in": dim_out, "dim_out": dim_out, "dropout_rate": params["dropout_rate"],
                      "activation": params["activation"]}
        if self.use_intra:
            self.intra_conv = IntraSO3ConvBlock(**intra_args)
        self.stride = params["stride"]
        self.skip_conv = nn.Conv2d(dim_in, dim_out, 1)
        self.norm = nn.InstanceNorm2d(dim_out, affine=False) if norm is None else norm(dim_out)
        self.relu = getattr(F, params["activation"])

    def forward(self, x, inter_idx, inter_w):
        skip_feature = x.feats
        skip_unit = bool(getattr(x, "is_occupancy", False)) or getattr(skip_feature, "_epn_unit", False)
        fused = None
        ic, ia = self.inter_conv, (self.intra_conv if self.use_intra else None)
        if ia is not None and not (self.training and (ic.dropout is not None or ia.dropout is not None)):
            # inter conv -> [norm + act applied inside the intra conv's operand load] -> intra conv -> norm + act
            occ = bool(getattr(x, "is_occupancy", False))
            with fwd_operands(None if occ else x.feats, occupancy=occ):
                inter_idx, inter_w, sample_idx, xr = ic.conv(x, inter_idx, inter_w)
            with fwd_operands(occupancy=True):   # the intra conv's operand is a normalised activation by construction
                fused = norm_intra(ic.norm, xr.feats, ic.relu, ia.conv)
            if fused is not None:
                x = sptk.SphericalPointCloud(xr.xyz, norm_act(ia.norm, fused, ia.relu), ia.conv.anchors)
            else:      # not a fusable combination: finish the inter block the ordinary way
                xr = sptk.SphericalPointCloud(xr.xyz, norm_act(ic.norm, xr.feats, ic.relu), xr.anchors)
                x = ia(xr)
        else:
            inter_idx, inter_w, sample_idx, x = self.inter_conv(x, inter_idx, inter_w)
            if self.use_intra:
                x = self.intra_conv(x)
        if self.stride > 1:
            if self.inter_conv.conv.lazy_sample:   # prefix sampling (pc/sample.py:64-67): the gather is a slice
                skip_feature = skip_feature[:, :, :sample_idx.shape[1]]
            else:
                skip_feature = L.batched_index_select(skip_feature, 2, sample_idx.long())
        # 1x1 skip conv = a BasicSO3Conv with kernel size 1: run it through the library's fp32-faithful
        # channel GEMM (cuDNN would silently use TF32), parameters stay in nn.Conv2d for checkpoint parity
        w = self.skip_conv.weight.view(self.skip_conv.out_channels, self.skip_conv.in_channels)
        with fwd_operands(occupancy=skip_unit):   # a strided slice of a normalised activation is one too
            skip_feature = sptk._BasicConvFn.apply(skip_feature.unsqueeze(2), w)
        # bias add, normalisation, activation and the residual add in one pass (see norm_act)
        feats = norm_act(self.norm, skip_feature, self.relu, residual=x.feats, bias=self.skip_conv.bias)
        x_out = sptk.SphericalPointCloud(x.xyz, feats, x.anchors)
        return inter_idx, inter_w, sample_idx, x_out

    def get_anchor(self):
        return torch.from_numpy(L.get_anchors())


class BasicSO3ConvBlock(nn.Module):
    """A list of conv blocks (base_so3conv.py:129-166)"""

    def __init__(self, params):
        super().__init__()
        self.blocks = nn.ModuleList()
        self.layer_types = []
        for param in params:
            if param["type"] == "intra_block":
                conv = IntraSO3ConvBlock(**param["args"])
            elif param["type"] == "inter_block":
                conv = InterSO3ConvBlock(**param["args"])
            elif param["type"] == "separable_block":
                conv = SeparableSO3ConvBlock(param["args"])
            else:
                raise ValueError("No such type of SO3Conv %s" % param["type"])
            self.layer_types.append(param["type"])
            self.blocks.append(conv)
        self.params = params

    def forward(self, x):
        inter_idx, inter_w = None, None
        for conv, param in zip(self.blocks, self.params):
            if param["type"] in ["inter", "inter_block", "separable_block"]:
                inter_idx, inter_w, _, x = conv(x, inter_idx, inter_w)
                if param["args"]["stride"] > 1:
                    inter_idx, inter_w = None, None
            elif param["type"] in ["intra_block"]:
                x = conv(x)
            else:
                raise ValueError("No such type of SO3Conv %s" % param["type"])
        return x

    def get_anchor(self):
        return torch.from_numpy(L.get_anchors())


# --------------------------------------------------------- backbone arithmetic
def backbone_params(input_num=1024, kanchor=60, dropout_rate=0.0,
                    mlps=((64, 64), (128, 128), (256, 256), (256,)), strides=(2, 2, 2, 2),
                    initial_radius_ratio=0.2, sampling_ratio=0.4, sampling_density=0.5,
                    kernel_multiplier=2, input_radius=1.0, sigma_ratio=0.5, xyz_pooling=None,
                    norm="BatchNorm2d", sigma_rule="double", scale_first_neighbor=False):
    """Layer hyper-parameters (per-layer radius / sigma / neighbour count / stride) of the three shipped
    backbones, i.e. the arithmetic of the reference's `build_model` functions:
      cls (cls_so3net_pn.py:41-150):  norm="BatchNorm2d", sigma doubles per block;
      reg (reg_so3net.py:54-171):     no norm key (InstanceNorm2d default), sigma doubles per block;
      inv (inv_so3net_pn.py:43-163):  no norm key, sigma multiplied by the block's stride (:98-100), first layer's
                                      neighbour count scaled by input_num/1024 (:112-113)."""
    strides = list(strides)
    na = kanchor
    if input_num > 1024:
        sampling_ratio /= (input_num / 1024)
        strides[0] = int(2 * (input_num / 1024))
    n_layer = len(mlps)
    stride_multipliers = [2 ** i for i in range(n_layer + 1)]
    num_centers = [int(input_num / m) for m in stride_multipliers]
    radius_ratio = [initial_radius_ratio * m ** sampling_density for m in stride_multipliers]
    radii = [r * input_radius for r in radius_ratio]
    weighted_sigma = [sigma_ratio * radii[0] ** 2]
    for i in range(len(strides)):
        weighted_sigma.append(weighted_sigma[i] * (2 if sigma_rule == "double" else strides[i]))
    backbone = []
    dim_in = 1
    for i, block in enumerate(mlps):
        block_param = []
        for j, dim_out in enumerate(block):
            lazy_sample = i != 0 or j != 0
            stride_conv = i == 0 or xyz_pooling != "stride"
            neighbor = int(sampling_ratio * num_centers[i] * radius_ratio[i] ** (1 / sampling_density))
            if scale_first_neighbor and i == 0 and j == 0:
                neighbor *= int(input_num / 1024)
            if j == 0:
                inter_stride = strides[i]
                nidx = i if i == 0 else i + 1
                if stride_conv:
                    neighbor *= 2
            else:
                inter_stride = 1
                nidx = i + 1
            args = {"dim_in": dim_in, "dim_out": dim_out, "kernel_size": 1, "stride": inter_stride,
                    "radius": radii[nidx], "sigma": weighted_sigma[nidx], "n_neighbor": neighbor,
                    "lazy_sample": lazy_sample, "dropout_rate": dropout_rate, "multiplier": kernel_multiplier,
                    "activation": "leaky_relu", "pooling": xyz_pooling, "kanchor": na}
            if norm is not None:
                args["norm"] = norm
            block_param.append({"type": "inter_block" if na != 60 else "separable_block", "args": args})
            dim_in = dim_out
        backbone.append(block_param)
    return backbone


def cls_backbone_params(input_num=1024, kanchor=60, dropout_rate=0.0, **kw):
    """ModelNet40 classification backbone (cls_so3net_pn.py:41-150)."""
    return backbone_params(input_num, kanchor, dropout_rate, **kw)


class SO3ConvBackbone(nn.Module):
    """preprocess_input + the list of BasicSO3ConvBlocks, i.e. ClsSO3ConvModel.forward minus the
    classification head (cls_so3net_pn.py:16-36)."""

    def __init__(self, backbone_params, na):
        super().__init__()
        self.backbone = nn.ModuleList([BasicSO3ConvBlock(bp) for bp in backbone_params])
        self.na_in = na

    def forward(self, x):
        x = preprocess_input(x, self.na_in, False)
        for block in self.backbone:
            x = block(x)
        return x
