"""Build + ctypes binding of libepn_b200.so (the C ABI declared in include/epn_b200.h).

The library is built IN-TREE (epn_pointcloud_b200/libepn_b200.so) by `build()` with
nvcc for sm_100a only.  There is no fallback: if the shared object is missing or a
call returns non-zero, a RuntimeError is raised.
"""
import ctypes
import glob
import os
import subprocess

_HERE = os.path.dirname(os.path.abspath(__file__))
_CSRC = os.path.join(_HERE, "csrc")
_SO = os.path.join(_HERE, "libepn_b200.so")
_HEADER = os.path.join(os.path.dirname(_HERE), "include", "epn_b200.h")

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-std=c++17",
    "-Xcompiler", "-fPIC,-fvisibility=hidden", "-shared",
]


def sources():
    return sorted(glob.glob(os.path.join(_CSRC, "*.cu")))


def _stale():
    if not os.path.exists(_SO):
        return True
    t = os.path.getmtime(_SO)
    deps = sources() + glob.glob(os.path.join(_CSRC, "*.cuh")) + [_HEADER]
    return any(os.path.getmtime(d) > t for d in deps)


def build(force=False, verbose=False):
    """nvcc -> epn_pointcloud_b200/libepn_b200.so (cross-compiles without a GPU).  Every .cu is compiled to
    an object under csrc/build/ (in parallel, re-used while newer than the source and every header), then linked."""
    if not force and not _stale():
        return _SO
    nvcc = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
    objdir = os.path.join(_CSRC, "build")
    os.makedirs(objdir, exist_ok=True)
    headers = glob.glob(os.path.join(_CSRC, "*.cuh")) + [_HEADER]
    hdr_t = max(os.path.getmtime(h) for h in headers)
    flags = [f for f in NVCC_FLAGS if f != "-shared"] + (["-Xptxas", "-v"] if verbose else [])
    jobs, objs = [], []
    for src in sources():
        obj = os.path.join(objdir, os.path.basename(src)[:-3] + ".o")
        objs.append(obj)
        if force or not os.path.exists(obj) or os.path.getmtime(obj) < max(os.path.getmtime(src), hdr_t):
            jobs.append((src, subprocess.Popen([nvcc] + flags + ["-c", "-o", obj, src])))
    failed = [src for src, j in jobs if j.wait() != 0]
    if failed:
        raise RuntimeError("nvcc failed for %s" % ", ".join(failed))
    subprocess.check_call([nvcc, "-gencode", "arch=compute_100a,code=sm_100a", "-shared", "-o", _SO] + objs)
    return _SO


c_f = ctypes.c_void_p   # device pointers travel as void*
c_i = ctypes.c_int
c_fl = ctypes.c_float
c_sz = ctypes.c_size_t
c_u64 = ctypes.c_ulonglong

# name -> (restype, argtypes); mirrors include/epn_b200.h one to one
SIGNATURES = {
    "epn_version": (c_i, []),
    "epn_last_error": (ctypes.c_char_p, []),
    "epn_device_supported": (c_i, []),
    "epn_launch_count": (ctypes.c_ulonglong, []),
    "epn_profile_enable": (None, [c_i]),
    "epn_profile_read": (c_i, [c_f, c_f, c_i]),
    "epn_ball_query_f32": (c_i, [c_f, c_f, c_f, c_i, c_i, c_i, c_fl, c_i, c_f]),
    "epn_fps_workspace_bytes": (c_sz, [c_i, c_i]),
    "epn_fps_f32": (c_i, [c_f, c_f, c_f, c_i, c_i, c_i, c_f]),
    "epn_gather_fwd_f32": (c_i, [c_f, c_f, c_f, c_i, c_i, c_i, c_i, c_f]),
    "epn_gather_bwd_f32": (c_i, [c_f, c_f, c_f, c_i, c_i, c_i, c_i, c_f]),
    "epn_zp_inter_fwd_f32": (c_i, [c_f] * 4 + [c_i] * 7 + [c_f]),
    "epn_zp_inter_bwd_f32": (c_i, [c_f] * 4 + [c_i] * 7 + [c_f]),
    "epn_zp_intra_fwd_f32": (c_i, [c_f] * 4 + [c_i] * 7 + [c_f]),
    "epn_zp_intra_bwd_f32": (c_i, [c_f] * 4 + [c_i] * 7 + [c_f]),
    "epn_inter_weights_f32": (c_i, [c_f] * 5 + [c_fl, c_f] + [c_i] * 6 + [c_f]),
    "epn_inter_group_fwd_f32": (c_i, [c_f] * 7 + [c_fl, c_f] + [c_i] * 7 + [c_f]),
    "epn_inter_group_bwd_f32": (c_i, [c_f] * 7 + [c_fl, c_f] + [c_i] * 7 + [c_f]),
    "epn_intra_group_fwd_f32": (c_i, [c_f] * 3 + [c_i] * 5 + [c_f]),
    "epn_intra_group_bwd_f32": (c_i, [c_f] * 3 + [c_i] * 5 + [c_f]),
    "epn_inter_so3conv_workspace_bytes": (c_sz, [c_i] * 9),
    "epn_inter_so3conv_grouped_bytes": (c_sz, [c_i] * 6),
    "epn_inter_so3conv_fwd_f32": (c_i, [c_f] * 6 + [c_fl, c_f, c_f, c_f, c_sz, c_f, c_sz, c_f] + [c_i] * 8 + [c_f]),
    "epn_inter_so3conv_bwd_f32": (c_i, [c_f] * 7 + [c_fl, c_f, c_f, c_f, c_f, c_sz, c_f, c_sz, c_u64] + [c_i] * 8 + [c_f]),
    "epn_intra_so3conv_workspace_bytes": (c_sz, [c_i] * 7),
    "epn_intra_so3conv_grouped_bytes": (c_sz, [c_i] * 5),
    "epn_intra_so3conv_fwd_f32": (c_i, [c_f] * 5 + [c_sz, c_f, c_sz, c_f] + [c_i] * 6 + [c_f]),
    "epn_intra_so3conv_bwd_f32": (c_i, [c_f] * 7 + [c_sz, c_f, c_sz, c_u64] + [c_i] * 6 + [c_f]),
    "epn_intra_so3conv_fwd_norm_f32": (c_i, [c_f] * 4 + [c_i, c_fl] + [c_f] * 4 + [c_sz, c_f, c_sz, c_f] + [c_i] * 6 + [c_f]),
    "epn_norm_stats_f32": (c_i, [c_f] * 3 + [c_sz] + [c_i] * 4 + [c_fl, c_f]),
    "epn_bn_track_f32": (c_i, [c_f] * 5 + [c_i, ctypes.c_longlong, c_fl, c_fl, c_f]),
    "epn_basic_conv_workspace_bytes": (c_sz, [c_i] * 4),
    "epn_basic_conv_fwd_f32": (c_i, [c_f] * 4 + [c_sz] + [c_i] * 4 + [c_f]),
    "epn_basic_conv_bwd_f32": (c_i, [c_f] * 6 + [c_sz] + [c_i] * 4 + [c_f]),
    "epn_norm_act_workspace_bytes": (c_sz, [c_i, c_i]),
    "epn_norm_act_fwd_f32": (c_i, [c_f] * 7 + [c_sz] + [c_i] * 4 + [c_fl, c_fl, c_f]),
    "epn_norm_act_bwd_f32": (c_i, [c_f] * 9 + [c_sz] + [c_i] * 4 + [c_fl, c_f]),
    "epn_set_gemm_backend": (None, [c_i]),
    "epn_set_slab_bytes": (None, [c_sz]),
    "epn_get_slab_bytes": (c_sz, []),
    "epn_get_gemm_backend": (c_i, []),
    "epn_set_fused_inter": (None, [c_i]),
    "epn_get_fused_inter": (c_i, []),
    "epn_set_fused_inter_bwd": (None, [c_i]),
    "epn_get_fused_inter_bwd": (c_i, []),
    "epn_set_forward_operands": (None, [c_i]),
    "epn_get_forward_operands": (c_i, []),
}

_lib = None


def lib():
    """Load the shared object (never builds implicitly on a GPU box: the .so ships in-tree)."""
    global _lib
    if _lib is None:
        if not os.path.exists(_SO):
            raise RuntimeError(
                "epn_pointcloud_b200: %s is missing -- run `python -c 'import __graft_entry__ as g; g.build()'` "
                "(there is no CPU or PyTorch fallback for the hot path)" % _SO)
        L = ctypes.CDLL(_SO)
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(L, name)  # AttributeError here == header/library mismatch
            fn.restype = res
            fn.argtypes = args
        _lib = L
    return _lib


def check(rc, what):
    if rc != 0:
        msg = lib().epn_last_error()
        raise RuntimeError("%s failed (rc=%d): %s" % (what, rc, msg.decode() if msg else "?"))
