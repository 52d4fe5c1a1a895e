"""Drop-in modules for the reference's `vgtk.so3conv` conv layers (boundary #2).

Same constructor signatures, forward return tuples, parameter/buffer names and shapes
(checkpoint compatible: `basic_conv.W [dim_out, dim_in*ks]`, `anchors`, `kernels`,
`intra_idx`) as vgtk/vgtk/so3conv/modules.py:21-55 (BasicSO3Conv), :125-174 (InterSO3Conv),
:177-200 (IntraSO3Conv); the container mirrors vgtk/vgtk/spconv/base.py:4-22.
Forward and backward of each layer are single calls into libepn_b200.so.
"""
import math

import torch
import torch.nn as nn

from . import functional as L
from . import ops

KERNEL_CONDENSE_RATIO = 0.7  # so3conv/modules.py:16


class PointSet:
    """point3d/base.py:15-58 (container part)"""

    def __init__(self, p):
        self._p = p

    @property
    def data(self):
        return self._p

    @property
    def n_batch(self):
        return self._p.shape[0]

    @property
    def n_point(self):
        return self._p.shape[-1]

    @property
    def device(self):
        return self._p.device


class SphericalPointCloud:
    """spconv/base.py:4-22.  `occupancy=(nb, np, na)` marks the network input whose features are
    the constant ones tensor (so3conv/functional.py:25-44): it is only materialised if somebody
    reads `.feats`; the first InterSO3Conv consumes the flag instead of 4*np*na bytes per cloud."""

    def __init__(self, xyz, feats, anchors, occupancy=None):
        self._xyz = PointSet(xyz)
        self._feats = feats
        self._anchors = anchors
        self._occupancy = occupancy

    @property
    def xyz(self):
        return self._xyz.data

    @property
    def feats(self):
        if self._feats is None and self._occupancy is not None:
            nb, npts, na = self._occupancy
            self._feats = torch.ones(nb, 1, npts, na, dtype=torch.float32, device=self.xyz.device)
        return self._feats

    @property
    def anchors(self):
        return self._anchors

    @property
    def is_occupancy(self):
        return self._occupancy is not None


class LazyInterW:
    """Stand-in for the reference's inter_w [b,p,na,ks,nn] return value.  Shipped models never read
    it; `.materialize()` produces the tensor, and passing the object back into
    InterSO3Conv.forward re-derives the same weights in registers."""

    def __init__(self, xyz, centers, idx, anchors, kernels, sigma):
        self.xyz, self.centers, self.idx = xyz, centers, idx
        self.anchors, self.kernels, self.sigma = anchors, kernels, sigma

    def materialize(self):
        return ops.inter_weights(self.xyz, self.centers, self.idx, self.anchors, self.kernels, self.sigma)

    @property
    def shape(self):
        b, p, nn_ = self.idx.shape
        return (b, p, self.anchors.shape[0], self.kernels.shape[0], nn_)


# ------------------------------------------------------------------ autograd
class _BasicConvFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, W):
        x = x.contiguous()
        W = W.contiguous()
        ctx.save_for_backward(x, W)
        return ops.basic_conv_fwd(x, W)

    @staticmethod
    def backward(ctx, dout):
        x, W = ctx.saved_tensors
        dx, dW = ops.basic_conv_bwd(dout.contiguous(), x, W, ctx.needs_input_grad[0], ctx.needs_input_grad[1])
        return dx, dW


class _InterSO3ConvFn(torch.autograd.Function):
    """Saves feats, W, the index/geometry tensors and -- memory permitting (ops.set_keep_grouped) -- the bf16
    operand tiles of the grouped tensor, from which backward takes dW; inter_w never exists, and without the
    kept tiles G is recomputed."""

    @staticmethod
    def forward(ctx, feats, W, xyz, centers, idx, anchors, kernels, sigma, training=True):
        W = W.contiguous()
        feats = None if feats is None else feats.contiguous()
        ctx.sigma = sigma
        ctx.has_feats = feats is not None
        ctx.save_for_backward(*( [feats] if feats is not None else [] ), W, xyz, centers, idx, anchors, kernels)
        # `training` = grad mode of the CALLER (inside Function.forward grad mode is always off): inference keeps nothing
        keep = bool(ctx.needs_input_grad[1]) and bool(training)
        res = ops.inter_so3conv_fwd(feats, xyz, centers, idx, anchors, kernels, sigma, W, keep_grouped=keep)
        out, ctx.grouped = res if keep else (res, None)
        return out

    @staticmethod
    def backward(ctx, dout):
        saved = list(ctx.saved_tensors)
        feats = saved.pop(0) if ctx.has_feats else None
        W, xyz, centers, idx, anchors, kernels = saved
        need_df = ctx.has_feats and ctx.needs_input_grad[0]
        dfeats, dW = ops.inter_so3conv_bwd(dout.contiguous(), feats, xyz, centers, idx, anchors, kernels, ctx.sigma, W,
                                           need_dfeats=need_df, need_dw=ctx.needs_input_grad[1], grouped=ctx.grouped)
        ctx.grouped = None
        return dfeats, dW, None, None, None, None, None, None, None


class _IntraSO3ConvFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, feats, W, intra_idx, training=True):
        feats = feats.contiguous()
        W = W.contiguous()
        ctx.save_for_backward(feats, W, intra_idx)
        keep = bool(ctx.needs_input_grad[1]) and bool(training)  # see _InterSO3ConvFn
        res = ops.intra_so3conv_fwd(feats, intra_idx, W, keep_grouped=keep)
        out, ctx.grouped = res if keep else (res, None)
        return out

    @staticmethod
    def backward(ctx, dout):
        feats, W, intra_idx = ctx.saved_tensors
        dfeats, dW = ops.intra_so3conv_bwd(dout.contiguous(), feats, intra_idx, W, ctx.needs_input_grad[0],
                                           ctx.needs_input_grad[1], grouped=ctx.grouped)
        ctx.grouped = None
        return dfeats, dW, None, None


# ------------------------------------------------------------------- modules
class BasicSO3Conv(nn.Module):
    """[b, c1, k, p, a] -> [b, c2, p, a]  (so3conv/modules.py:21-55)"""

    def __init__(self, dim_in, dim_out, kernel_size, debug=False):
        super().__init__()
        self.dim_in = dim_in
        self.dim_out = dim_out
        self.kernel_size = kernel_size
        if debug:
            self.register_buffer("W", torch.ones(dim_out, dim_in * kernel_size))
        else:
            W = torch.empty(dim_out, dim_in, kernel_size)
            nn.init.xavier_normal_(W, gain=nn.init.calculate_gain("relu"))
            self.register_parameter("W", nn.Parameter(W.view(dim_out, dim_in * kernel_size)))

    def forward(self, x):
        return _BasicConvFn.apply(x, self.W)


class InterSO3Conv(nn.Module):
    """[b, c1, p1, a] -> [b, c2, p2, a]  (so3conv/modules.py:125-174)"""

    def __init__(self, dim_in, dim_out, kernel_size, stride, radius, sigma, n_neighbor,
                 lazy_sample=True, pooling=None, kanchor=60):
        super().__init__()
        kernels = L.get_sphereical_kernel_points_from_ply(KERNEL_CONDENSE_RATIO * radius, kernel_size)
        anchors = L.get_anchors(kanchor)
        self.dim_in = dim_in
        self.dim_out = dim_out
        self.kernel_size = kernels.shape[0]
        self.stride = stride
        self.radius = radius
        self.sigma = sigma
        self.n_neighbor = n_neighbor
        self.lazy_sample = lazy_sample
        self.pooling = pooling
        self.basic_conv = BasicSO3Conv(dim_in, dim_out, self.kernel_size)
        self.register_buffer("anchors", torch.from_numpy(anchors))
        self.register_buffer("kernels", torch.from_numpy(kernels))

    def forward(self, x, inter_idx=None, inter_w=None):
        xyz = x.xyz
        occupancy = getattr(x, "is_occupancy", False) and self.dim_in == 1
        feats = None if occupancy else x.feats
        if self.pooling is not None and self.stride > 1 and self.dim_in > 1:
            raise NotImplementedError("pooling=%r is unused by every shipped model" % (self.pooling,))
        W = self.basic_conv.W
        if inter_idx is None:
            # sampling + ball query (so3conv/functional.py:150-152 -> spconv/functional.py:412-421)
            n_sample = math.ceil(xyz.shape[2] / self.stride)
            sample_idx, new_xyz = L.furthest_sample(xyz, n_sample, self.lazy_sample)
            inter_idx = L.ball_query_index(new_xyz, xyz, self.radius, self.n_neighbor)
            inter_w = LazyInterW(xyz, new_xyz, inter_idx, self.anchors, self.kernels, self.sigma)
        else:
            sample_idx, new_xyz = None, xyz
            inter_idx = inter_idx.int().contiguous()
            if inter_w is None:
                inter_w = LazyInterW(xyz, new_xyz, inter_idx, self.anchors, self.kernels, self.sigma)
        if isinstance(inter_w, LazyInterW):
            out = _InterSO3ConvFn.apply(feats, W, inter_w.xyz, inter_w.centers, inter_idx, inter_w.anchors,
                                        inter_w.kernels, inter_w.sigma, torch.is_grad_enabled())
        else:  # caller supplied a materialised weight tensor: honour it (unfused op-surface path)
            grouped = L.inter_zpconv_grouping_naive(inter_idx, inter_w, x.feats)
            out = self.basic_conv(grouped)
        return inter_idx, inter_w, sample_idx, SphericalPointCloud(new_xyz, out, self.anchors)


class IntraSO3Conv(nn.Module):
    """Only defined for the 60-anchor group (so3conv/modules.py:177-200)."""

    def __init__(self, dim_in, dim_out):
        super().__init__()
        anchors = L.get_anchors()
        intra_idx = L.get_intra_idx()
        self.dim_in = dim_in
        self.dim_out = dim_out
        self.kernel_size = intra_idx.shape[1]
        self.basic_conv = BasicSO3Conv(dim_in, dim_out, self.kernel_size)
        self.register_buffer("anchors", torch.from_numpy(anchors))
        self.register_buffer("intra_idx", torch.from_numpy(intra_idx).long())
        self.register_buffer("_intra_idx32", torch.from_numpy(intra_idx).int().contiguous(), persistent=False)
        # the int32 copy the kernels read follows `intra_idx` when a checkpoint overwrites it
        self.register_load_state_dict_post_hook(IntraSO3Conv._sync_idx32)

    @staticmethod
    def _sync_idx32(module, incompatible_keys):
        module._intra_idx32.copy_(module.intra_idx.to(torch.int32))

    def forward(self, x):
        feats = _IntraSO3ConvFn.apply(x.feats, self.basic_conv.W, self._intra_idx32, torch.is_grad_enabled())
        return SphericalPointCloud(x.xyz, feats, self.anchors)
