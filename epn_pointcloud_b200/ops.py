"""Tensor-level bindings of the C ABI (boundary #1 of SURVEY.md section 8b).

Every function takes/returns CUDA torch tensors in the reference's layouts, allocates
the outputs with torch (the library itself never allocates), launches on torch's
current stream and raises RuntimeError on any failure.  The three namespaces
`grouping`, `gathering`, `zpconv` at the bottom expose exactly the names, argument
order and return shapes of the reference's pybind modules `vgtk.cuda.grouping`
(vgtk/vgtk/cuda/grouping_cuda.cpp:176-181), `vgtk.cuda.gathering`
(gathering_cuda.cpp:61-65) and `vgtk.cuda.zpconv` (zpconv_cuda.cpp:112-118).
"""
import ctypes
import contextlib
import types

import torch

from . import _lib


def _require_cuda(*tensors):
    for t in tensors:
        if t is None:
            continue
        if not t.is_cuda:  # same contract as CHECK_CUDA, grouping_cuda.cpp:66-68
            raise RuntimeError("epn_pointcloud_b200: tensor must be a CUDA tensor (no CPU fallback exists)")
        if not t.is_contiguous():
            raise RuntimeError("epn_pointcloud_b200: tensor must be contiguous")


def _f32(t):
    return t if t.dtype == torch.float32 else t.float()


def _i32(t):
    return t if t.dtype == torch.int32 else t.int()


def _p(t):
    return None if t is None else t.data_ptr()


def _stream():
    return torch.cuda.current_stream().cuda_stream


_ws_cache = {}


def _workspace(nbytes, device):
    """One grow-only scratch buffer per (device, stream); 256-B aligned by the torch allocator."""
    key = (device.index, torch.cuda.current_stream(device).cuda_stream)
    buf = _ws_cache.get(key)
    if buf is None or buf.numel() < nbytes:
        buf = torch.empty(max(int(nbytes), 256), dtype=torch.uint8, device=device)
        _ws_cache[key] = buf
    return buf


# ------------------------------------------------------------------ index ops
def ball_query(new_xyz, xyz, radius, nsample):
    """new_xyz [b,3,m], xyz [b,3,n] -> idx int32 [b,m,nsample] (grouping_cuda.cpp:71-86)."""
    new_xyz, xyz = _f32(new_xyz), _f32(xyz)
    _require_cuda(new_xyz, xyz)
    b, _, m = new_xyz.shape
    n = xyz.shape[2]
    idx = torch.empty(b, m, nsample, dtype=torch.int32, device=xyz.device)
    with torch.cuda.device(xyz.device):
        _lib.check(_lib.lib().epn_ball_query_f32(_p(new_xyz), _p(xyz), _p(idx), b, n, m, float(radius),
                                                 int(nsample), _stream()), "epn_ball_query_f32")
    return idx


def furthest_point_sampling(xyz, m):
    """xyz [b,3,n] -> idx int32 [b,m] (grouping_cuda.cpp:160-174)."""
    xyz = _f32(xyz)
    _require_cuda(xyz)
    b, _, n = xyz.shape
    idx = torch.empty(b, m, dtype=torch.int32, device=xyz.device)
    L = _lib.lib()
    with torch.cuda.device(xyz.device):
        wsb = L.epn_fps_workspace_bytes(b, n)
        ws = _workspace(wsb, xyz.device) if wsb else None
        _lib.check(L.epn_fps_f32(_p(xyz), _p(ws), _p(idx), b, n, int(m), _stream()), "epn_fps_f32")
    return idx


def gather_points_forward(points, idx):
    """points [b,c,n], idx int32 [b,m] -> float32 [b,c,m] (gathering_cuda.cpp:29-43)."""
    points, idx = _f32(points), _i32(idx)
    _require_cuda(points, idx)
    b, c, n = points.shape
    m = idx.shape[1]
    out = torch.empty(b, c, m, dtype=torch.float32, device=points.device)
    with torch.cuda.device(points.device):
        _lib.check(_lib.lib().epn_gather_fwd_f32(_p(points), _p(idx), _p(out), b, c, n, m, _stream()),
                   "epn_gather_fwd_f32")
    return out


def gather_points_backward(grad_out, idx, npoint):
    """grad_out [b,c,m], idx [b,m] -> [b,c,npoint] (gathering_cuda.cpp:46-59)."""
    grad_out, idx = _f32(grad_out), _i32(idx)
    _require_cuda(grad_out, idx)
    b, c, m = grad_out.shape
    out = torch.zeros(b, c, int(npoint), dtype=torch.float32, device=grad_out.device)
    with torch.cuda.device(grad_out.device):
        _lib.check(_lib.lib().epn_gather_bwd_f32(_p(grad_out), _p(idx), _p(out), b, c, int(npoint), m, _stream()),
                   "epn_gather_bwd_f32")
    return out


# ------------------------------------------------------------- zpconv surface
def inter_zpconv_forward(inter_idx, inter_w, feats):
    """idx/w [b,np,na,ks,ann], feats [b,c,nq,na] -> [b,c,ks,np,na] (zpconv_cuda.cpp:41-58)."""
    inter_idx, inter_w, feats = _i32(inter_idx), _f32(inter_w), _f32(feats)
    _require_cuda(inter_idx, inter_w, feats)
    b, np_, na, ks, ann = inter_idx.shape
    c, nq = feats.shape[1], feats.shape[2]
    out = torch.empty(b, c, ks, np_, na, dtype=torch.float32, device=feats.device)
    with torch.cuda.device(feats.device):
        _lib.check(_lib.lib().epn_zp_inter_fwd_f32(_p(inter_idx), _p(inter_w), _p(feats), _p(out), b, c, nq, np_, na,
                                                   ks, ann, _stream()), "epn_zp_inter_fwd_f32")
    return out


def inter_zpconv_backward(inter_idx, inter_w, grad, npoint):
    """grad [b,c,ks,np,na] -> [b,c,npoint,na] (zpconv_cuda.cpp:60-77)."""
    inter_idx, inter_w, grad = _i32(inter_idx), _f32(inter_w), _f32(grad)
    _require_cuda(inter_idx, inter_w, grad)
    b, np_, na, ks, ann = inter_idx.shape
    c = grad.shape[1]
    out = torch.zeros(b, c, int(npoint), na, dtype=torch.float32, device=grad.device)
    with torch.cuda.device(grad.device):
        _lib.check(_lib.lib().epn_zp_inter_bwd_f32(_p(inter_idx), _p(inter_w), _p(grad), _p(out), b, c, int(npoint),
                                                   np_, na, ks, ann, _stream()), "epn_zp_inter_bwd_f32")
    return out


def intra_zpconv_forward(intra_idx, intra_w, feats):
    """idx [na_out,ann], w [na_out,ks,ann], feats [b,c,np,na_in] -> [b,c,ks,np,na_out] (zpconv_cuda.cpp:79-95)."""
    intra_idx, intra_w, feats = _i32(intra_idx), _f32(intra_w), _f32(feats)
    _require_cuda(intra_idx, intra_w, feats)
    na_out, ann = intra_idx.shape
    ks = intra_w.shape[1]
    b, c, np_, na_in = feats.shape
    out = torch.empty(b, c, ks, np_, na_out, dtype=torch.float32, device=feats.device)
    with torch.cuda.device(feats.device):
        _lib.check(_lib.lib().epn_zp_intra_fwd_f32(_p(intra_idx), _p(intra_w), _p(feats), _p(out), b, c, np_, na_in,
                                                   na_out, ks, ann, _stream()), "epn_zp_intra_fwd_f32")
    return out


def intra_zpconv_backward(intra_idx, intra_w, grad, anchor_in):
    """grad [b,c,ks,np,na_out] -> [b,c,np,anchor_in] (zpconv_cuda.cpp:97-110)."""
    intra_idx, intra_w, grad = _i32(intra_idx), _f32(intra_w), _f32(grad)
    _require_cuda(intra_idx, intra_w, grad)
    na_out, ann = intra_idx.shape
    ks = intra_w.shape[1]
    b, c, _, np_, _ = grad.shape
    out = torch.zeros(b, c, np_, int(anchor_in), dtype=torch.float32, device=grad.device)
    with torch.cuda.device(grad.device):
        _lib.check(_lib.lib().epn_zp_intra_bwd_f32(_p(intra_idx), _p(intra_w), _p(grad), _p(out), b, c, np_,
                                                   int(anchor_in), na_out, ks, ann, _stream()), "epn_zp_intra_bwd_f32")
    return out


# ----------------------------------------------------- live-path grouping stages
def inter_weights(xyz, centers, idx, anchors, kernels, sigma):
    """-> inter_w [b,p,na,ks,nn] (so3conv/functional.py:180-218)."""
    _require_cuda(xyz, centers, idx, anchors, kernels)
    b, _, p_in = xyz.shape
    p, nn = idx.shape[1], idx.shape[2]
    na, ks = anchors.shape[0], kernels.shape[0]
    w = torch.empty(b, p, na, ks, nn, dtype=torch.float32, device=xyz.device)
    with torch.cuda.device(xyz.device):
        _lib.check(_lib.lib().epn_inter_weights_f32(_p(xyz), _p(centers), _p(idx), _p(anchors), _p(kernels),
                                                    float(sigma), _p(w), b, p_in, p, nn, na, ks, _stream()),
                   "epn_inter_weights_f32")
    return w


def inter_group_fwd(feats, idx, inter_w=None, geom=None):
    """feats [b,c,p_in,na] (or None = occupancy ones, c=1), idx [b,p,nn], and either inter_w
    [b,p,na,ks,nn] or geom=(xyz, centers, anchors, kernels, sigma) -> [b,c,ks,p,na]."""
    _require_cuda(feats, idx, inter_w)
    b, p, nn = idx.shape
    if inter_w is not None:
        na, ks = inter_w.shape[2], inter_w.shape[3]
        xyz = centers = anchors = kernels = None
        sigma = 1.0
    else:
        xyz, centers, anchors, kernels, sigma = geom
        _require_cuda(xyz, centers, anchors, kernels)
        na, ks = anchors.shape[0], kernels.shape[0]
    if feats is not None:
        c, p_in = feats.shape[1], feats.shape[2]
    else:
        c, p_in = 1, xyz.shape[2]
    out = torch.empty(b, c, ks, p, na, dtype=torch.float32, device=idx.device)
    with torch.cuda.device(idx.device):
        _lib.check(_lib.lib().epn_inter_group_fwd_f32(_p(feats), _p(idx), _p(inter_w), _p(xyz), _p(centers),
                                                      _p(anchors), _p(kernels), float(sigma), _p(out), b, c, p_in, p,
                                                      nn, na, ks, _stream()), "epn_inter_group_fwd_f32")
    return out


def inter_group_bwd(dout, idx, p_in, inter_w=None, geom=None):
    """dout [b,c,ks,p,na] -> dfeats [b,c,p_in,na]."""
    _require_cuda(dout, idx, inter_w)
    b, c, ks, p, na = dout.shape
    nn = idx.shape[2]
    if inter_w is not None:
        xyz = centers = anchors = kernels = None
        sigma = 1.0
    else:
        xyz, centers, anchors, kernels, sigma = geom
        _require_cuda(xyz, centers, anchors, kernels)
    dfeats = torch.zeros(b, c, int(p_in), na, dtype=torch.float32, device=dout.device)
    with torch.cuda.device(dout.device):
        _lib.check(_lib.lib().epn_inter_group_bwd_f32(_p(dout), _p(idx), _p(inter_w), _p(xyz), _p(centers),
                                                      _p(anchors), _p(kernels), float(sigma), _p(dfeats), b, c,
                                                      int(p_in), p, nn, na, ks, _stream()), "epn_inter_group_bwd_f32")
    return dfeats


def intra_group_fwd(feats, intra_idx):
    """feats [b,c,p,na], intra_idx int32 [na,kn] -> [b,c,kn,p,na] (so3conv/functional.py:221-268)."""
    _require_cuda(feats, intra_idx)
    b, c, p, na = feats.shape
    kn = intra_idx.shape[1]
    out = torch.empty(b, c, kn, p, na, dtype=torch.float32, device=feats.device)
    with torch.cuda.device(feats.device):
        _lib.check(_lib.lib().epn_intra_group_fwd_f32(_p(feats), _p(intra_idx), _p(out), b, c, p, na, kn, _stream()),
                   "epn_intra_group_fwd_f32")
    return out


def intra_group_bwd(dout, intra_idx):
    """dout [b,c,kn,p,na] -> dfeats [b,c,p,na]; columns of intra_idx must be permutations."""
    _require_cuda(dout, intra_idx)
    b, c, kn, p, na = dout.shape
    dfeats = torch.empty(b, c, p, na, dtype=torch.float32, device=dout.device)
    with torch.cuda.device(dout.device):
        _lib.check(_lib.lib().epn_intra_group_bwd_f32(_p(dout), _p(intra_idx), _p(dfeats), b, c, p, na, kn, _stream()),
                   "epn_intra_group_bwd_f32")
    return dfeats


# -------------------------------------------------------------------- fused convs
def basic_conv_fwd(x, W):
    """x [b,c,ks,p,na], W [co, c*ks] -> [b,co,p,na] (so3conv/modules.py:48-55)."""
    _require_cuda(x, W)
    b, c, ks, p, na = x.shape
    co = W.shape[0]
    out = torch.empty(b, co, p, na, dtype=torch.float32, device=x.device)
    L = _lib.lib()
    with torch.cuda.device(x.device):
        wsb = L.epn_basic_conv_workspace_bytes(b, c * ks, co, p * na)
        ws = _workspace(wsb, x.device)
        _lib.check(L.epn_basic_conv_fwd_f32(_p(x), _p(W), _p(out), _p(ws), wsb, b, c * ks, co, p * na, _stream()),
                   "epn_basic_conv_fwd_f32")
    return out


def basic_conv_bwd(dout, x, W, need_dx=True, need_dw=True):
    _require_cuda(dout, x, W)
    b, c, ks, p, na = x.shape
    co = W.shape[0]
    dx = torch.empty_like(x) if need_dx else None
    dW = torch.empty_like(W) if need_dw else None
    L = _lib.lib()
    with torch.cuda.device(x.device):
        wsb = L.epn_basic_conv_workspace_bytes(b, c * ks, co, p * na)
        ws = _workspace(wsb, x.device)
        _lib.check(L.epn_basic_conv_bwd_f32(_p(dout), _p(x), _p(W), _p(dx), _p(dW), _p(ws), wsb, b, c * ks, co, p * na,
                                            _stream()), "epn_basic_conv_bwd_f32")
    return dx, dW


def _grouped_buffer(nbytes, device, keep_grouped):
    """uint8 buffer for the operand tiles a training forward keeps, or None (see set_keep_grouped)."""
    if not keep_grouped or nbytes == 0:
        return None
    mode = keep_grouped_mode()
    if mode == "off":
        return None
    if mode == "auto":
        free, _ = torch.cuda.mem_get_info(device)
        free += torch.cuda.memory_reserved(device) - torch.cuda.memory_allocated(device)
        if nbytes > _KEEP_FRACTION * free:
            return None
    return torch.empty(int(nbytes), dtype=torch.uint8, device=device)


class _Grouped:
    """The operand tiles a training forward kept (uint8 device buffer) together with the layout word the
    library returned for them (K order + slab plan); the backward gets both back, so the pair stays consistent
    whatever happens to the process-wide knobs in between."""

    __slots__ = ("buf", "layout")

    def __init__(self, buf, layout):
        self.buf, self.layout = buf, int(layout)

    @staticmethod
    def wrap(buf, layout):
        return None if buf is None else _Grouped(buf, layout)

    @staticmethod
    def args(g):
        """(pointer, bytes, layout) arguments of a backward call."""
        if g is None:
            return None, 0, 0
        return g.buf.data_ptr(), g.buf.numel(), g.layout

    def numel(self):
        return self.buf.numel()


_KEEP_FRACTION = 0.25
_keep_mode = None


def keep_grouped_mode():
    """'auto' (default): a forward under autograd keeps the grouped operand tiles for dW when they take less
    than a quarter of the free device memory; 'on': always; 'off': never (dW recomputes the grouping).
    Initial value from EPN_KEEP_GROUPED."""
    global _keep_mode
    if _keep_mode is None:
        import os
        v = os.environ.get("EPN_KEEP_GROUPED", "auto").lower()
        _keep_mode = {"1": "on", "0": "off"}.get(v, v)
        if _keep_mode not in ("auto", "on", "off"):
            raise RuntimeError("EPN_KEEP_GROUPED must be auto, on or off")
    return _keep_mode


def set_keep_grouped(mode):
    global _keep_mode
    if mode not in ("auto", "on", "off"):
        raise ValueError("mode must be 'auto', 'on' or 'off'")
    _keep_mode = mode


def inter_so3conv_fwd(feats, xyz, centers, idx, anchors, kernels, sigma, W, keep_grouped=False):
    """Fused InterSO3Conv minus sampling/ball query -> [b,c_out,p,na].  feats None = occupancy ones.
    keep_grouped=True returns (out, grouped) with grouped = the kept operand tiles (or None) for
    inter_so3conv_bwd(..., grouped=grouped)."""
    _require_cuda(feats, xyz, centers, idx, anchors, kernels, W)
    b, _, p_in = xyz.shape
    p, nn = idx.shape[1], idx.shape[2]
    na, ks = anchors.shape[0], kernels.shape[0]
    c_in = 1 if feats is None else feats.shape[1]
    c_out = W.shape[0]
    if W.shape[1] != c_in * ks:
        raise RuntimeError("W must be [c_out, c_in*ks] = [%d, %d], got %s" % (c_out, c_in * ks, tuple(W.shape)))
    out = torch.empty(b, c_out, p, na, dtype=torch.float32, device=xyz.device)
    L = _lib.lib()
    with torch.cuda.device(xyz.device):
        wsb = L.epn_inter_so3conv_workspace_bytes(b, c_in, c_out, p_in, p, nn, na, ks, 0)
        ws = _workspace(wsb, xyz.device)
        grouped = _grouped_buffer(L.epn_inter_so3conv_grouped_bytes(b, c_in, p, nn, na, ks) if keep_grouped else 0,
                                  xyz.device, keep_grouped)
        layout = ctypes.c_ulonglong(0)
        _lib.check(L.epn_inter_so3conv_fwd_f32(_p(feats), _p(xyz), _p(centers), _p(idx), _p(anchors), _p(kernels),
                                               float(sigma), _p(W), _p(out), _p(ws), wsb, _p(grouped),
                                               0 if grouped is None else grouped.numel(), ctypes.addressof(layout),
                                               b, c_in, c_out, p_in, p, nn, na, ks, _stream()),
                   "epn_inter_so3conv_fwd_f32")
    return (out, _Grouped.wrap(grouped, layout.value)) if keep_grouped else out


def inter_so3conv_bwd(dout, feats, xyz, centers, idx, anchors, kernels, sigma, W, need_dfeats=True, need_dw=True,
                      grouped=None):
    _require_cuda(dout, feats, xyz, centers, idx, anchors, kernels, W)
    b, _, p_in = xyz.shape
    p, nn = idx.shape[1], idx.shape[2]
    na, ks = anchors.shape[0], kernels.shape[0]
    c_in = 1 if feats is None else feats.shape[1]
    c_out = W.shape[0]
    dfeats = (torch.empty(b, c_in, p_in, na, dtype=torch.float32, device=dout.device)
              if (need_dfeats and feats is not None) else None)
    dW = torch.empty_like(W) if need_dw else None
    L = _lib.lib()
    with torch.cuda.device(dout.device):
        wsb = L.epn_inter_so3conv_workspace_bytes(b, c_in, c_out, p_in, p, nn, na, ks, 1)
        ws = _workspace(wsb, dout.device)
        _lib.check(L.epn_inter_so3conv_bwd_f32(_p(dout), _p(feats), _p(xyz), _p(centers), _p(idx), _p(anchors),
                                               _p(kernels), float(sigma), _p(W), _p(dfeats), _p(dW), _p(ws), wsb,
                                               *_Grouped.args(grouped), b,
                                               c_in, c_out, p_in, p, nn, na, ks, _stream()),
                   "epn_inter_so3conv_bwd_f32")
    return dfeats, dW


def intra_so3conv_fwd(feats, intra_idx, W, keep_grouped=False):
    """Fused IntraSO3Conv: feats [b,c,p,na], intra_idx int32 [na,kn], W [co, c*kn] -> [b,co,p,na]
    (keep_grouped: as inter_so3conv_fwd)."""
    _require_cuda(feats, intra_idx, W)
    b, c_in, p, na = feats.shape
    kn = intra_idx.shape[1]
    c_out = W.shape[0]
    if W.shape[1] != c_in * kn:
        raise RuntimeError("W must be [c_out, c_in*kn]")
    out = torch.empty(b, c_out, p, na, dtype=torch.float32, device=feats.device)
    L = _lib.lib()
    with torch.cuda.device(feats.device):
        wsb = L.epn_intra_so3conv_workspace_bytes(b, c_in, c_out, p, na, kn, 0)
        ws = _workspace(wsb, feats.device)
        grouped = _grouped_buffer(L.epn_intra_so3conv_grouped_bytes(b, c_in, p, na, kn) if keep_grouped else 0,
                                  feats.device, keep_grouped)
        layout = ctypes.c_ulonglong(0)
        _lib.check(L.epn_intra_so3conv_fwd_f32(_p(feats), _p(intra_idx), _p(W), _p(out), _p(ws), wsb, _p(grouped),
                                               0 if grouped is None else grouped.numel(), ctypes.addressof(layout),
                                               b, c_in, c_out, p, na, kn, _stream()), "epn_intra_so3conv_fwd_f32")
    return (out, _Grouped.wrap(grouped, layout.value)) if keep_grouped else out


def intra_so3conv_fwd_norm(x, stats, gamma, beta, mode, slope, intra_idx, W, keep_grouped=False):
    """IntraSO3Conv applied to leaky_relu(norm(x) * gamma + beta, slope) WITHOUT materialising that activation: x
    [b,c,p,na] is the raw output of the preceding conv, stats [2,G] its (mean, rstd) (norm_stats, or the running
    statistics of an evaluation-mode BatchNorm with mode 1); the normalisation is applied while the operand tiles are
    built (epn_intra_so3conv_fwd_norm_f32).  Returns None instead of raising when the shape is outside the tile
    routes, so the caller can run the unfused sequence."""
    _require_cuda(x, stats, gamma, beta, intra_idx, W)
    b, c_in, p, na = x.shape
    kn = intra_idx.shape[1]
    c_out = W.shape[0]
    if W.shape[1] != c_in * kn:
        raise RuntimeError("W must be [c_out, c_in*kn]")
    L = _lib.lib()
    if L.epn_get_gemm_backend() != 0 or na != 60 or kn != 12 or p % 4 != 0:
        return None
    out = torch.empty(b, c_out, p, na, dtype=torch.float32, device=x.device)
    with torch.cuda.device(x.device):
        wsb = L.epn_intra_so3conv_workspace_bytes(b, c_in, c_out, p, na, kn, 0)
        ws = _workspace(wsb, x.device)
        grouped = _grouped_buffer(L.epn_intra_so3conv_grouped_bytes(b, c_in, p, na, kn) if keep_grouped else 0,
                                  x.device, keep_grouped)
        if keep_grouped and grouped is None:
            return None   # a backward without kept tiles would have to re-gather the (never materialised) activation
        layout = ctypes.c_ulonglong(0)
        _lib.check(L.epn_intra_so3conv_fwd_norm_f32(_p(x), _p(stats), _p(gamma), _p(beta), int(mode), float(slope), _p(intra_idx),
                                                    _p(W), _p(out), _p(ws), wsb, _p(grouped),
                                                    0 if grouped is None else grouped.numel(), ctypes.addressof(layout),
                                                    b, c_in, c_out, p, na, kn, _stream()), "epn_intra_so3conv_fwd_norm_f32")
    return (out, _Grouped.wrap(grouped, layout.value)) if keep_grouped else out


def intra_so3conv_bwd(dout, feats, intra_idx, W, need_dfeats=True, need_dw=True, grouped=None):
    _require_cuda(dout, feats, intra_idx, W)
    b, c_in, p, na = feats.shape
    kn = intra_idx.shape[1]
    c_out = W.shape[0]
    dfeats = torch.empty_like(feats) if need_dfeats else None
    dW = torch.empty_like(W) if need_dw else None
    L = _lib.lib()
    with torch.cuda.device(feats.device):
        wsb = L.epn_intra_so3conv_workspace_bytes(b, c_in, c_out, p, na, kn, 1)
        ws = _workspace(wsb, feats.device)
        _lib.check(L.epn_intra_so3conv_bwd_f32(_p(dout), _p(feats), _p(intra_idx), _p(W), _p(dfeats), _p(dW), _p(ws),
                                               wsb, *_Grouped.args(grouped), b, c_in,
                                               c_out, p, na, kn, _stream()),
                   "epn_intra_so3conv_bwd_f32")
    return dfeats, dW


def norm_act_fwd(x, gamma, beta, mode, eps, slope, residual=None, stats=None):
    """x [b,c,p,a] -> (y, stats[2,G]); mode 0 = InstanceNorm2d(affine=False), 1 = BatchNorm2d (batch statistics),
    2 = BatchNorm2d in evaluation mode (`stats` [2,c] = per-channel (mean, 1/sqrt(var+eps)) given by the caller),
    followed by leaky_relu(slope)  (SPConvNets/utils/base_so3conv.py:43,55-57,107,119-125); `residual` (same shape)
    is added to the result in the same pass (the skip connection, base_so3conv.py:209-211)."""
    _require_cuda(x, gamma, beta, residual, stats)
    b, c = x.shape[0], x.shape[1]
    n = x[0, 0].numel()
    y = torch.empty_like(x)
    if mode == 2:
        if stats is None or tuple(stats.shape) != (2, c) or stats.dtype != torch.float32 or not stats.is_contiguous():
            raise ValueError("mode 2 needs stats [2, c] float32 contiguous")
    else:
        stats = torch.empty(2, b * c if mode == 0 else c, dtype=torch.float32, device=x.device)
    L = _lib.lib()
    with torch.cuda.device(x.device):
        wsb = L.epn_norm_act_workspace_bytes(b, c)
        ws = _workspace(wsb, x.device)
        _lib.check(L.epn_norm_act_fwd_f32(_p(x), _p(gamma), _p(beta), _p(residual), _p(y), _p(stats), _p(ws), wsb, b, c, n, int(mode),
                                          float(eps), float(slope), _stream()), "epn_norm_act_fwd_f32")
    return y, stats


def norm_stats(x, mode, eps):
    """(mean, rstd) [2,G] of x [b,c,p,a]: G = b*c (mode 0, InstanceNorm2d) or c (mode 1, BatchNorm2d batch statistics)."""
    _require_cuda(x)
    b, c = x.shape[0], x.shape[1]
    n = x[0, 0].numel()
    stats = torch.empty(2, b * c if mode == 0 else c, dtype=torch.float32, device=x.device)
    L = _lib.lib()
    with torch.cuda.device(x.device):
        wsb = L.epn_norm_act_workspace_bytes(b, c)
        ws = _workspace(wsb, x.device)
        _lib.check(L.epn_norm_stats_f32(_p(x), _p(stats), _p(ws), wsb, b, c, n, int(mode), float(eps), _stream()), "epn_norm_stats_f32")
    return stats


def bn_track(stats, bias, running_mean, running_var, num_batches_tracked, count, momentum, eps):
    """In-place running-statistics update of a training-mode BatchNorm from the batch (mean, rstd) `stats` [2,c]
    (epn_bn_track_f32: one launch instead of eight elementwise ones); momentum None = cumulative average."""
    _require_cuda(stats, bias, running_mean, running_var, num_batches_tracked)
    if num_batches_tracked.dtype != torch.int64:
        raise RuntimeError("num_batches_tracked must be int64")
    c = running_mean.numel()
    with torch.cuda.device(stats.device):
        _lib.check(_lib.lib().epn_bn_track_f32(_p(stats), _p(bias), _p(running_mean), _p(running_var), _p(num_batches_tracked), c,
                                               int(count), -1.0 if momentum is None else float(momentum), float(eps), _stream()),
                   "epn_bn_track_f32")


def norm_act_bwd(dy, x, gamma, beta, stats, mode, slope, need_affine_grads=True):
    _require_cuda(dy, x, gamma, beta, stats)
    b, c = x.shape[0], x.shape[1]
    n = x[0, 0].numel()
    dx = torch.empty_like(x)
    dgamma = dbeta = None
    if mode == 1 and gamma is not None and need_affine_grads:
        dgamma, dbeta = torch.empty_like(gamma), torch.empty_like(gamma)
    L = _lib.lib()
    with torch.cuda.device(x.device):
        wsb = L.epn_norm_act_workspace_bytes(b, c)
        ws = _workspace(wsb, x.device)
        _lib.check(L.epn_norm_act_bwd_f32(_p(dy), _p(x), _p(gamma), _p(beta), _p(stats), _p(dx), _p(dgamma), _p(dbeta),
                                          _p(ws), wsb, b, c, n, int(mode), float(slope), _stream()), "epn_norm_act_bwd_f32")
    return dx, dgamma, dbeta


def set_gemm_backend(name):
    """'umma' (tcgen05 tensor cores, default) or 'simt' (fp32 cross-check path)."""
    _lib.lib().epn_set_gemm_backend({"umma": 0, "simt": 1}[name])


def set_fused_inter(on):
    """True (default): InterSO3Conv forward runs as ONE fused kernel; False: grouping kernel + GEMM kernel."""
    _lib.lib().epn_set_fused_inter(1 if on else 0)


def set_fused_inter_bwd(on):
    """1 / True: the data gradient of InterSO3Conv (rows of <= 16 slots) runs as ONE fused kernel; 2: rows of 17..32
    slots too; 0 / False: GEMM + scatter."""
    _lib.lib().epn_set_fused_inter_bwd(int(on))


@contextlib.contextmanager
def forward_operands(fmt):
    """Operand format of the forward GEMMs issued by THIS thread inside the block: 'bf16' (default: fp32's range,
    ~2^-18 per operand) or 'f16' (22 significand bits at the same speed, fp16's range: for inputs known to be
    normalised activations; honoured only by forwards that keep no operand tiles).  See include/epn_b200.h."""
    L = _lib.lib()
    old = L.epn_get_forward_operands()
    L.epn_set_forward_operands({"bf16": 0, "f16": 1}[fmt])
    try:
        yield
    finally:
        L.epn_set_forward_operands(old)


# ------------------------------------------ drop-in namespaces for vgtk.cuda.*
grouping = types.SimpleNamespace(ball_query=ball_query, furthest_point_sampling=furthest_point_sampling)
gathering = types.SimpleNamespace(gather_points_forward=gather_points_forward,
                                  gather_points_backward=gather_points_backward)
zpconv = types.SimpleNamespace(inter_zpconv_forward=inter_zpconv_forward,
                               inter_zpconv_backward=inter_zpconv_backward,
                               intra_zpconv_forward=intra_zpconv_forward,
                               intra_zpconv_backward=intra_zpconv_backward)
