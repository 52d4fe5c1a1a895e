"""Host-side mirror of the reference's functional layer for the hot path.

Same names, argument meaning and return tuples as
  vgtk/vgtk/pc/sample.py:46-77          (group_nd, ball_query_index, furthest_sample[_index])
  vgtk/vgtk/utils.py:25-27              (batch_gather)
  vgtk/vgtk/spconv/functional.py:83-128,340-421   (shadow helpers, Gathering, ball_query,
                                         batched_index_select, inter_zpconv_grouping_naive/_ball,
                                         InterZPConvGrouping, IntraZPConvGrouping)
  vgtk/vgtk/so3conv/functional.py:25-44,86-96,118-299  (occupancy features, kernel points,
                                         inter/intra_so3conv_grouping, anchors)
but every tensor op on the path is one call into libepn_b200.so -- no torch.gather /
einsum / index_select chains, no shadow-point copies, no CPU fallback.
"""
import math
import os

import numpy as np
import torch

from . import ops

_DATA = os.path.join(os.path.dirname(os.path.abspath(__file__)), "data", "so3_constants.npz")
_consts = None


def _constants():
    """60 icosahedral rotations, 60x12 intra index and raw kernel point sets.  Generated once by
    oracle/make_constants.py from the reference's import-time initialisation
    (so3conv/functional.py:274-278, functional/rotation.py:236-343)."""
    global _consts
    if _consts is None:
        with np.load(_DATA) as d:
            _consts = {k: d[k] for k in d.files}
    return _consts


# ------------------------------------------------------------------ constants
def select_anchor(anchors, k):
    """so3conv/functional.py:281-289"""
    if k == 1:
        return anchors[29][None]
    if k == 20:
        return anchors[::3]
    if k == 40:
        return anchors.reshape(20, 3, 3, 3)[:, :2].reshape(-1, 3, 3)
    return anchors


def get_anchors(k=60):
    return np.ascontiguousarray(select_anchor(_constants()["anchors"], k))


def get_intra_idx():
    return _constants()["intra_idx"]


def get_sphereical_kernel_points_from_ply(radius, kernel_size):
    """so3conv/functional.py:86-96: kpsphere{24,30,66}.ply rescaled so the max norm equals `radius`."""
    assert 0 < kernel_size <= 3
    pts = _constants()["kpsphere%d" % {1: 24, 2: 30, 3: 66}[kernel_size]].astype("float32")
    r = np.sqrt((pts ** 2).sum(1).max())
    return pts * radius / r


def get_occupancy_features(pc, n_anchor, use_center=False):
    """so3conv/functional.py:25-44 (xyz-only input): ones [nb,1,np,na]."""
    nb, npts, nd = pc.shape
    if nd != 3:
        raise NotImplementedError("normals branch is broken upstream (so3conv/functional.py:35-36)")
    feats = torch.ones(nb, 1, npts, n_anchor, dtype=torch.float32, device=pc.device)
    if use_center:
        feats[:, :, 0, :] = 0.0
    return feats


# ---------------------------------------------------------------- index helpers
def batch_gather(x, idx, dim=2):
    """utils.py:25-27"""
    return ops.gather_points_forward(x.contiguous(), idx.int().contiguous())


def group_nd(pc, idx):
    """pc [b,c,n] x idx [b,m1(,m2..)] -> [b,c,m1(,m2..)]  (pc/sample.py:46-50)"""
    b = idx.shape[0]
    out = batch_gather(pc, idx.reshape(b, -1).contiguous())
    return out.view(b, -1, *idx.shape[1:])


def ball_query_index(query_points, support_points, radius, n_sample):
    """pc/sample.py:54-59"""
    return ops.ball_query(query_points.contiguous(), support_points.contiguous(), radius, n_sample)


def furthest_sample_index(pc, n_sample, lazy_sample):
    """pc/sample.py:63-72"""
    if pc.shape[2] == n_sample or lazy_sample:
        return torch.arange(n_sample, dtype=torch.int32, device=pc.device).view(1, -1).expand(pc.shape[0], -1).contiguous()
    return ops.furthest_point_sampling(pc.contiguous(), n_sample)


def furthest_sample(pc, n_sample, lazy_sample=True):
    """pc/sample.py:75-77"""
    idx = furthest_sample_index(pc, n_sample, lazy_sample)
    return idx, group_nd(pc, idx)


def ball_query(query_points, support_points, radius, n_sample, support_feats=None):
    """spconv/functional.py:340-349 -> (idx, grouped_xyz[, grouped_feats]).  The reference appends a
    shadow point at 1e4 before gathering; no index ever addresses it (the CUDA op never emits n),
    so the gather runs on the un-padded cloud with identical results."""
    idx = ball_query_index(query_points, support_points, radius, n_sample)
    if support_feats is None:
        return idx, group_nd(support_points, idx)
    return idx, group_nd(support_points, idx), group_nd(support_feats, idx)


def batched_index_select(input, dim, index):
    """spconv/functional.py:361-369.  Off the conv hot path (skip connections / heads); differentiable."""
    for ii in range(1, input.dim()):
        if ii != dim:
            index = index.unsqueeze(ii)
    expanse = list(input.shape)
    expanse[0] = -1
    expanse[dim] = -1
    return torch.gather(input, dim, index.expand(expanse))


class Gathering(torch.autograd.Function):
    """spconv/functional.py:101-128"""

    @staticmethod
    def forward(ctx, points, idx):
        ctx.save_for_backward(idx)
        ctx.npoint = points.size(2)
        return ops.gather_points_forward(points.contiguous(), idx.int().contiguous())

    @staticmethod
    def backward(ctx, grad):
        (idx,) = ctx.saved_tensors
        return ops.gather_points_backward(grad.contiguous(), idx.int().contiguous(), ctx.npoint), None


# ------------------------------------------------------------ zpconv surface
class InterZPConvGrouping(torch.autograd.Function):
    """spconv/functional.py:313-334"""

    @staticmethod
    def forward(ctx, inter_idx, inter_w, feats):
        ctx.save_for_backward(inter_idx, inter_w)
        ctx.npoint = feats.size(2)
        return ops.inter_zpconv_forward(inter_idx.contiguous(), inter_w.contiguous(), feats.contiguous())

    @staticmethod
    def backward(ctx, grad):
        inter_idx, inter_w = ctx.saved_tensors
        return None, None, ops.inter_zpconv_backward(inter_idx.contiguous(), inter_w.contiguous(), grad.contiguous(),
                                                     ctx.npoint)


class IntraZPConvGrouping(torch.autograd.Function):
    """spconv/functional.py:210-237"""

    @staticmethod
    def forward(ctx, intra_idx, intra_w, feats):
        ctx.save_for_backward(intra_idx, intra_w)
        ctx.anchor_in = feats.size(3)
        return ops.intra_zpconv_forward(intra_idx.contiguous(), intra_w.contiguous(), feats.contiguous())

    @staticmethod
    def backward(ctx, grad):
        intra_idx, intra_w = ctx.saved_tensors
        return None, None, ops.intra_zpconv_backward(intra_idx.contiguous(), intra_w.contiguous(), grad.contiguous(),
                                                     ctx.anchor_in)


def intra_zpconv_grouping(intra_idx, intra_w, feats):
    return IntraZPConvGrouping.apply(intra_idx, intra_w, feats)


# -------------------------------------------------------- live grouping stages
class _InterGroupW(torch.autograd.Function):
    """out[b,c,k,p,a] = sum_n feats[b,c,idx[b,p,n],a] * inter_w[b,p,a,k,n]; grad flows to feats only
    (inter_w carries no grad in the reference either)."""

    @staticmethod
    def forward(ctx, inter_idx, inter_w, feats):
        ctx.save_for_backward(inter_idx, inter_w)
        ctx.p_in = feats.size(2)
        return ops.inter_group_fwd(feats.contiguous(), inter_idx, inter_w=inter_w)

    @staticmethod
    def backward(ctx, grad):
        inter_idx, inter_w = ctx.saved_tensors
        return None, None, ops.inter_group_bwd(grad.contiguous(), inter_idx, ctx.p_in, inter_w=inter_w)


def inter_zpconv_grouping_naive(inter_idx, inter_w, feats):
    """spconv/functional.py:372-390.  feats may carry the reference's shadow row (p_in+1 rows): it is
    never addressed, so it is accepted and ignored."""
    return _InterGroupW.apply(inter_idx.int().contiguous(), inter_w.contiguous(), feats)


inter_so3conv_feat_grouping = inter_zpconv_grouping_naive


class _IntraGroup(torch.autograd.Function):
    @staticmethod
    def forward(ctx, intra_idx, feats):
        ctx.save_for_backward(intra_idx)
        return ops.intra_group_fwd(feats.contiguous(), intra_idx)

    @staticmethod
    def backward(ctx, grad):
        (intra_idx,) = ctx.saved_tensors
        return None, ops.intra_group_bwd(grad.contiguous(), intra_idx)


def intra_so3conv_grouping(intra_idx, feature):
    """so3conv/functional.py:221-268 -> [nb, c_in, pnn, np, na]"""
    return _IntraGroup.apply(intra_idx.int().contiguous(), feature)


def inter_zpconv_grouping_ball(xyz, stride, radius, n_neighbor, lazy_sample=True):
    """spconv/functional.py:412-421 -> (grouped_xyz - centre, ball_idx, sample_idx, sample_xyz)"""
    n_sample = math.ceil(xyz.shape[2] / stride)
    idx, sample_xyz = furthest_sample(xyz, n_sample, lazy_sample)
    ball_idx, grouped_xyz = ball_query(sample_xyz, xyz, radius, n_neighbor)
    grouped_xyz = grouped_xyz - sample_xyz.unsqueeze(3)
    return grouped_xyz, ball_idx, idx, sample_xyz


def inter_so3conv_grouping_anchor(grouped_xyz, anchors, kernels, sigma, interpolate="linear"):
    """so3conv/functional.py:180-218 on an explicit grouped_xyz [b,3,p,nn] (already centre-relative).
    Off the fused path (which derives the weights in registers); provided for op parity."""
    if interpolate != "linear":
        raise NotImplementedError("kernel function %s is not implemented!" % interpolate)
    b, _, p, nn = grouped_xyz.shape
    # neighbours become a private 'cloud' of p*nn points per batch with zero centres
    xyz = grouped_xyz.reshape(b, 3, p * nn).contiguous()
    centers = torch.zeros(b, 3, p, dtype=torch.float32, device=xyz.device)
    idx = torch.arange(p * nn, dtype=torch.int32, device=xyz.device).view(1, p, nn).expand(b, -1, -1).contiguous()
    return ops.inter_weights(xyz, centers, idx, anchors.contiguous(), kernels.contiguous(), sigma)


def inter_so3conv_grouping(xyz, feats, stride, n_neighbor, anchors, kernels, radius, sigma,
                           inter_idx=None, inter_w=None, lazy_sample=True, radius_expansion=1.0, pooling=None):
    """so3conv/functional.py:118-178 -> (inter_idx, inter_w, new_xyz, grouped_feats, sample_idx).
    Unfused op-surface form (materialises inter_w like the reference); InterSO3Conv.forward uses the
    fused kernel instead."""
    if pooling is not None and stride > 1 and feats.shape[1] > 1:
        raise NotImplementedError("pooling=%r: unused by every shipped model (xyz_pooling=None)" % (pooling,))
    if inter_idx is None:
        n_sample = math.ceil(xyz.shape[2] / stride)
        sample_idx, new_xyz = furthest_sample(xyz, n_sample, lazy_sample)
        inter_idx = ball_query_index(new_xyz, xyz, radius * radius_expansion, n_neighbor)
        inter_w = ops.inter_weights(xyz.contiguous(), new_xyz, inter_idx, anchors, kernels, sigma)
    else:
        sample_idx = None
        new_xyz = xyz
    new_feats = inter_so3conv_feat_grouping(inter_idx, inter_w, feats)
    return inter_idx, inter_w, new_xyz, new_feats, sample_idx
