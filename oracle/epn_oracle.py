"""ctypes front-end of the C oracle (oracle/epn_oracle.c) on CPU torch tensors.

TEST INFRASTRUCTURE ONLY: imported by tests/, __graft_entry__.smoke() and the
cpu_baseline / --impl reference legs of bench.py -- never by the product
package `epn_pointcloud_b200`.

Function names and argument order follow the reference's pybind surface
(vgtk/vgtk/cuda/grouping_cuda.cpp:71-181, gathering_cuda.cpp:29-65,
zpconv_cuda.cpp:41-118) so the harness can serve them as `vgtk.cuda.*`.
"""
import ctypes
import os
import subprocess

import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "libepn_oracle.so")
_SRC = os.path.join(_HERE, "epn_oracle.c")
_lib = None


def build(force=False):
    """gcc -O2, contraction OFF so that only the explicit fmaf() calls fuse."""
    if force or not os.path.exists(_SO) or os.path.getmtime(_SO) < os.path.getmtime(_SRC):
        subprocess.check_call(["gcc", "-O2", "-ffp-contract=off", "-fPIC", "-shared",
                               "-fvisibility=hidden", "-o", _SO, _SRC, "-lm"])
    return _SO


def lib():
    global _lib
    if _lib is None:
        _lib = ctypes.CDLL(build())
    return _lib


def _f(t):
    assert t.dtype == torch.float32 and t.is_contiguous() and t.device.type == "cpu"
    return ctypes.c_void_p(t.data_ptr())


def _i(t):
    assert t.dtype == torch.int32 and t.is_contiguous() and t.device.type == "cpu"
    return ctypes.c_void_p(t.data_ptr())


def _cf(t):
    return t.detach().to(torch.float32).contiguous()


def _ci(t):
    return t.detach().to(torch.int32).contiguous()


# ------------------------------------------------------------- live native ops
def ball_query(new_xyz, xyz, radius, nsample):
    new_xyz, xyz = _cf(new_xyz), _cf(xyz)
    b, _, m = new_xyz.shape
    n = xyz.shape[2]
    idx = torch.zeros(b, m, nsample, dtype=torch.int32)
    lib().epn_oracle_ball_query(b, n, m, ctypes.c_float(radius), nsample, _f(new_xyz), _f(xyz), _i(idx))
    return idx


def furthest_point_sampling(xyz, m):
    xyz = _cf(xyz)
    b, _, n = xyz.shape
    idx = torch.zeros(b, m, dtype=torch.int32)
    lib().epn_oracle_fps(b, n, m, _f(xyz), _i(idx))
    return idx


def gather_points_forward(points, idx):
    points, idx = _cf(points), _ci(idx)
    b, c, n = points.shape
    m = idx.shape[1]
    out = torch.empty(b, c, m, dtype=torch.float32)
    lib().epn_oracle_gather_fwd(b, c, n, m, _f(points), _i(idx), _f(out))
    return out


def gather_points_backward(grad_out, idx, npoint):
    grad_out, idx = _cf(grad_out), _ci(idx)
    b, c, m = grad_out.shape
    out = torch.empty(b, c, npoint, dtype=torch.float32)
    lib().epn_oracle_gather_bwd(b, c, npoint, m, _f(grad_out), _i(idx), _f(out))
    return out


# ------------------------------------------------------------------ conv stages
def inter_weights(xyz, centers, idx, anchors, kernels, sigma):
    xyz, centers, idx, anchors, kernels = _cf(xyz), _cf(centers), _ci(idx), _cf(anchors), _cf(kernels)
    b, _, p_in = xyz.shape
    p, nn = idx.shape[1], idx.shape[2]
    na, ks = anchors.shape[0], kernels.shape[0]
    w = torch.empty(b, p, na, ks, nn, dtype=torch.float32)
    lib().epn_oracle_inter_weights(b, p_in, p, nn, na, ks, _f(xyz), _f(centers), _i(idx), _f(anchors),
                                   _f(kernels), ctypes.c_float(sigma), _f(w))
    return w


def inter_group_fwd(idx, inter_w, feats):
    idx, inter_w, feats = _ci(idx), _cf(inter_w), _cf(feats)
    b, c, p_in, na = feats.shape
    p, nn = idx.shape[1], idx.shape[2]
    ks = inter_w.shape[3]
    out = torch.empty(b, c, ks, p, na, dtype=torch.float32)
    lib().epn_oracle_inter_group_fwd(b, c, p_in, p, nn, na, ks, _f(feats), _i(idx), _f(inter_w), _f(out))
    return out


def inter_group_bwd(idx, inter_w, dout, p_in):
    idx, inter_w, dout = _ci(idx), _cf(inter_w), _cf(dout)
    b, c, ks, p, na = dout.shape
    nn = idx.shape[2]
    out = torch.empty(b, c, p_in, na, dtype=torch.float32)
    lib().epn_oracle_inter_group_bwd(b, c, p_in, p, nn, na, ks, _f(dout), _i(idx), _f(inter_w), _f(out))
    return out


def intra_group_fwd(intra_idx, feats):
    intra_idx, feats = _ci(intra_idx), _cf(feats)
    b, c, p, na = feats.shape
    kn = intra_idx.shape[1]
    out = torch.empty(b, c, kn, p, na, dtype=torch.float32)
    lib().epn_oracle_intra_group_fwd(b, c, p, na, kn, _f(feats), _i(intra_idx), _f(out))
    return out


def intra_group_bwd(intra_idx, dout):
    intra_idx, dout = _ci(intra_idx), _cf(dout)
    b, c, kn, p, na = dout.shape
    out = torch.empty(b, c, p, na, dtype=torch.float32)
    lib().epn_oracle_intra_group_bwd(b, c, p, na, kn, _f(dout), _i(intra_idx), _f(out))
    return out


def basic_conv(x, W):
    """x [B,C,KS,P,A], W [CO, C*KS] -> [B,CO,P,A]"""
    x, W = _cf(x), _cf(W)
    b, c, ks, p, na = x.shape
    co = W.shape[0]
    out = torch.empty(b, co, p, na, dtype=torch.float32)
    lib().epn_oracle_basic_conv(b, c * ks, co, p * na, _f(x), _f(W), _f(out))
    return out


# -------------------------------------------------------------- zpconv surface
def zp_inter_forward(nbr, w, feats):
    nbr, w, feats = _ci(nbr), _cf(w), _cf(feats)
    b, np_, na, ks, ann = nbr.shape
    c, nq = feats.shape[1], feats.shape[2]
    out = torch.empty(b, c, ks, np_, na, dtype=torch.float32)
    lib().epn_oracle_zp_inter_fwd(b, c, nq, np_, na, ks, ann, _i(nbr), _f(w), _f(feats), _f(out))
    return out


def zp_inter_backward(nbr, w, dout, npoint):
    nbr, w, dout = _ci(nbr), _cf(w), _cf(dout)
    b, np_, na, ks, ann = nbr.shape
    c = dout.shape[1]
    out = torch.empty(b, c, npoint, na, dtype=torch.float32)
    lib().epn_oracle_zp_inter_bwd(b, c, npoint, np_, na, ks, ann, _i(nbr), _f(w), _f(dout), _f(out))
    return out


def zp_intra_forward(nbr, w, feats):
    nbr, w, feats = _ci(nbr), _cf(w), _cf(feats)
    na_out, ann = nbr.shape
    ks = w.shape[1]
    b, c, np_, na_in = feats.shape
    out = torch.empty(b, c, ks, np_, na_out, dtype=torch.float32)
    lib().epn_oracle_zp_intra_fwd(b, c, np_, na_in, na_out, ks, ann, _i(nbr), _f(w), _f(feats), _f(out))
    return out


def zp_intra_backward(nbr, w, dout, anchor_in):
    nbr, w, dout = _ci(nbr), _cf(w), _cf(dout)
    na_out, ann = nbr.shape
    ks = w.shape[1]
    b, c, _, np_, _ = dout.shape
    out = torch.empty(b, c, np_, anchor_in, dtype=torch.float32)
    lib().epn_oracle_zp_intra_bwd(b, c, np_, anchor_in, na_out, ks, ann, _i(nbr), _f(w), _f(dout), _f(out))
    return out


# ------------------------------------------------------- composed conv layers
def inter_so3conv(xyz, feats, W, anchors, kernels, stride, n_neighbor, radius, sigma, lazy_sample=True):
    """InterSO3Conv.forward restated on the C oracle
    (vgtk/vgtk/so3conv/modules.py:157-174 -> so3conv/functional.py:118-178 ->
    spconv/functional.py:412-421).  Returns (inter_idx, inter_w, new_xyz, out, sample_idx)."""
    xyz = _cf(xyz)
    p_in = xyz.shape[2]
    n_sample = -(-p_in // stride)
    if p_in == n_sample or lazy_sample:  # vgtk/vgtk/pc/sample.py:63-67
        sample_idx = torch.arange(n_sample, dtype=torch.int32).expand(xyz.shape[0], -1).contiguous()
    else:
        sample_idx = furthest_point_sampling(xyz, n_sample)
    new_xyz = gather_points_forward(xyz, sample_idx)
    idx = ball_query(new_xyz, xyz, radius, n_neighbor)
    w = inter_weights(xyz, new_xyz, idx, anchors, kernels, sigma)
    g = inter_group_fwd(idx, w, feats)
    return idx, w, new_xyz, basic_conv(g, W), sample_idx


def intra_so3conv(feats, W, intra_idx):
    """IntraSO3Conv.forward restated (vgtk/vgtk/so3conv/modules.py:197-200)."""
    return basic_conv(intra_group_fwd(intra_idx, feats), W)
