"""Stub-import harness that runs the UNMODIFIED reference Python (on CPU, or on the GPU with the
reference's own CUDA extensions from oracle/_ref).

TEST INFRASTRUCTURE ONLY (see oracle/README.md).  Used in the build container
(where /root/reference is mounted) by oracle/make_golden.py and
oracle/make_constants.py to pin the oracle and to generate tests/golden/*, and by
bench.py's reference arms (`--impl reference`, `cpu_baseline`, `reference_gpu`) -- the
only places that may execute anything under oracle/.  On the GPU box /root/reference
does not exist: there the harness reads the reference's Python from baseline/_ref/, a
git-ignored verbatim copy made in the build container by oracle/vendor_ref.py (the
reference is a Python package with an unbuildable setup.py -- it compiles its CUDA
extensions at install time -- so the copy stands in for `pip install --target
baseline/_ref`).  Nothing from the reference is committed to the repository.

What is stubbed:
  * trimesh (==3.2.0 upstream, not installed): `load()` returns an object with
    `faces`, `face_normals`, `face_adjacency`, `fix_normals()` computed from a
    binary-PLY read of vgtk/data/anchors/sphere12.ply.  Consumed at
    vgtk/vgtk/functional/rotation.py:236-243,117-139.
  * plyfile: `PlyData.read()` for the ASCII kernel-point files
    (vgtk/vgtk/pc/io.py:6-9).
  * open3d, parse, colour, imageio: empty modules (data loading only).
  * vgtk.cuda.{grouping,gathering,zpconv}: the compiled extensions.  The three
    live ops are served by the C oracle (oracle/epn_oracle.c); the zpconv
    entry points by the C oracle's zp_* functions.
"""
import importlib
import os
import struct
import sys
import types

import numpy as np
import torch

_REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
VENDORED_ROOT = os.path.join(_REPO, "baseline", "_ref")


def _find_root():
    for cand in (os.environ.get("EPN_REFERENCE_ROOT"), "/root/reference", VENDORED_ROOT):
        if cand and os.path.isdir(os.path.join(cand, "vgtk", "vgtk")) and os.path.isdir(os.path.join(cand, "SPConvNets")):
            return cand
    return os.environ.get("EPN_REFERENCE_ROOT", "/root/reference")


REFERENCE_ROOT = _find_root()


def available():
    """True when the reference's Python tree can be imported (build container, or baseline/_ref on the GPU box)."""
    return os.path.isdir(os.path.join(REFERENCE_ROOT, "vgtk", "vgtk"))


# ----------------------------------------------------------------- PLY readers
def _read_ply_header(f):
    lines = []
    while True:
        line = f.readline().decode("ascii").strip()
        lines.append(line)
        if line == "end_header":
            break
    return lines


def read_binary_icosahedron(path):
    """sphere12.ply: vertex = 3 f32 + 4 u8, face = u8 n, n i32, u8 m, m f32, 4 u8."""
    with open(path, "rb") as f:
        hdr = _read_ply_header(f)
        nv = int([l for l in hdr if l.startswith("element vertex")][0].split()[-1])
        nf = int([l for l in hdr if l.startswith("element face")][0].split()[-1])
        verts = np.zeros((nv, 3), dtype=np.float32)
        for i in range(nv):
            rec = f.read(16)
            verts[i] = struct.unpack("<3f", rec[:12])
        faces = np.zeros((nf, 3), dtype=np.int64)
        for i in range(nf):
            (n,) = struct.unpack("<B", f.read(1))
            faces[i] = struct.unpack("<%di" % n, f.read(4 * n))
            (m,) = struct.unpack("<B", f.read(1))
            f.read(4 * m + 4)
    return verts, faces


def read_ascii_vertices(path):
    with open(path, "rb") as f:
        hdr = _read_ply_header(f)
        nv = int([l for l in hdr if l.startswith("element vertex")][0].split()[-1])
        pts = np.zeros((nv, 3), dtype=np.float32)
        for i in range(nv):
            pts[i] = [float(t) for t in f.readline().decode("ascii").split()[:3]]
    return pts


# ------------------------------------------------------------------ stub mods
class _Mesh:
    def __init__(self, path):
        v, f = read_binary_icosahedron(path)
        self.vertices = v.astype(np.float64)
        self.faces = f
        tri = self.vertices[f]
        nrm = np.cross(tri[:, 1] - tri[:, 0], tri[:, 2] - tri[:, 0])
        self.face_normals = nrm / np.linalg.norm(nrm, axis=1, keepdims=True)
        # edge -> faces, edges ordered by (lo, hi), face pairs ascending
        edges = {}
        for fi, face in enumerate(f):
            for e in range(3):
                a, b = int(face[e]), int(face[(e + 1) % 3])
                edges.setdefault((min(a, b), max(a, b)), []).append(fi)
        self.face_adjacency = np.array([sorted(edges[k]) for k in sorted(edges)], dtype=np.int64)

    def fix_normals(self):
        # all 20 faces of sphere12.ply are already outward wound
        assert (np.einsum("ij,ij->i", self.face_normals, self.vertices[self.faces].mean(1)) > 0).all()


def _make_plyfile():
    mod = types.ModuleType("plyfile")

    class PlyData(dict):
        @staticmethod
        def read(path):
            pts = read_ascii_vertices(path)
            d = PlyData()
            d["vertex"] = {"x": pts[:, 0], "y": pts[:, 1], "z": pts[:, 2]}
            return d

    class PlyElement:  # pragma: no cover - never used on the hot path
        pass

    mod.PlyData = PlyData
    mod.PlyElement = PlyElement
    return mod


def _make_cuda_stubs():
    """vgtk.cuda.* : CPU tensors are served by the C oracle; CUDA tensors by the REFERENCE's own CUDA
    extensions compiled into oracle/_ref (oracle/build_ref.py) -- never by this repository's library."""
    from oracle import epn_oracle as O
    from oracle import build_ref

    ref_ext = {}

    def ext(name):
        if name not in ref_ext:
            ref_ext[name] = build_ref.load_ref(name)
            if ref_ext[name] is None:
                raise RuntimeError("oracle/_ref/vgtk_ref_%s.so not built (python -m oracle.build_ref)" % name)
        return ref_ext[name]

    def dual(ext_name, fn_name, cpu_fn):
        def call(*a):
            if any(isinstance(t, torch.Tensor) and t.is_cuda for t in a):
                with torch.cuda.device(next(t.device for t in a if isinstance(t, torch.Tensor) and t.is_cuda)):
                    return getattr(ext(ext_name), fn_name)(*a)
            return cpu_fn(*a)
        return call

    pkg = types.ModuleType("vgtk.cuda")
    pkg.__path__ = []
    grouping = types.ModuleType("vgtk.cuda.grouping")
    gathering = types.ModuleType("vgtk.cuda.gathering")
    zpconv = types.ModuleType("vgtk.cuda.zpconv")

    grouping.ball_query = dual("grouping", "ball_query", lambda new_xyz, xyz, radius, nsample: O.ball_query(new_xyz, xyz, radius, nsample))
    grouping.furthest_point_sampling = dual("grouping", "furthest_point_sampling", lambda xyz, m: O.furthest_point_sampling(xyz, m))
    gathering.gather_points_forward = dual("gathering", "gather_points_forward", lambda pts, idx: O.gather_points_forward(pts, idx))
    gathering.gather_points_backward = dual("gathering", "gather_points_backward", lambda g, idx, n: O.gather_points_backward(g, idx, n))
    zpconv.inter_zpconv_forward = dual("zpconv", "inter_zpconv_forward", O.zp_inter_forward)
    zpconv.inter_zpconv_backward = dual("zpconv", "inter_zpconv_backward", O.zp_inter_backward)
    zpconv.intra_zpconv_forward = dual("zpconv", "intra_zpconv_forward", O.zp_intra_forward)
    zpconv.intra_zpconv_backward = dual("zpconv", "intra_zpconv_backward", O.zp_intra_backward)
    pkg.grouping, pkg.gathering, pkg.zpconv = grouping, gathering, zpconv
    return {"vgtk.cuda": pkg, "vgtk.cuda.grouping": grouping,
            "vgtk.cuda.gathering": gathering, "vgtk.cuda.zpconv": zpconv}


_loaded = {}


def load_reference():
    """Import the reference's `vgtk` (and return it).  Idempotent."""
    if "vgtk" in _loaded:
        return _loaded["vgtk"]
    if not os.path.isdir(REFERENCE_ROOT):
        raise RuntimeError("reference tree not present at %s" % REFERENCE_ROOT)
    trimesh = types.ModuleType("trimesh")
    trimesh.load = lambda path, *a, **k: _Mesh(path)
    sys.modules.setdefault("trimesh", trimesh)
    sys.modules.setdefault("plyfile", _make_plyfile())
    for name in ("open3d", "parse", "colour", "imageio"):
        sys.modules.setdefault(name, types.ModuleType(name))
    sys.modules["parse"].parse = lambda *a, **k: None
    for k, v in _make_cuda_stubs().items():
        sys.modules[k] = v
    sys.path.insert(0, os.path.join(REFERENCE_ROOT, "vgtk"))
    sys.path.insert(0, REFERENCE_ROOT)
    argv, sys.argv = sys.argv, [sys.argv[0]]
    try:
        vgtk = importlib.import_module("vgtk")
    finally:
        sys.argv = argv
    # the stubbed sub-package must be reachable as an attribute too
    vgtk.cuda = sys.modules["vgtk.cuda"]
    _loaded["vgtk"] = vgtk
    return vgtk


def load_spconvnets(gpu_ops=None):
    """Import SPConvNets.utils.base_so3conv + the three model builders.  (`gpu_ops` is informational: the
    vgtk.cuda stubs pick the reference's CUDA extensions or the C oracle per call from the tensors' device.)"""
    load_reference()
    if "M" in _loaded:
        return _loaded["M"]
    argv, sys.argv = sys.argv, [sys.argv[0], "experiment"]
    try:
        M = importlib.import_module("SPConvNets.utils.base_so3conv")
    finally:
        sys.argv = argv
    _loaded["M"] = M
    return M


def cls_opt(input_num=1024, kanchor=60, device="cpu"):
    """The slice of the reference's `opt` namespace its build_model functions read
    (SPConvNets/options.py:9-103, SPConvNets/models/cls_so3net_pn.py:58-66)."""
    o = types.SimpleNamespace()
    o.device = device
    o.model = types.SimpleNamespace(input_num=input_num, dropout_rate=0.0, kpconv=False, kanchor=kanchor, flag="max",
                                    search_radius=0.4, representation="quat")
    o.train_loss = types.SimpleNamespace(temperature=3.0)
    return o


def build_cls_model(input_num=1024, kanchor=60):
    """The reference's classification network, built by ITS build_model (default init, caller seeds)."""
    import contextlib
    import io
    load_spconvnets()
    mod = importlib.import_module("SPConvNets.models.cls_so3net_pn")
    with contextlib.redirect_stdout(io.StringIO()):
        return mod.build_model(cls_opt(input_num, kanchor))
