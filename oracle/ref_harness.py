"""Stub-import harness that runs the UNMODIFIED reference Python on CPU.

TEST INFRASTRUCTURE ONLY (see oracle/README.md).  Used in the build container
(where /root/reference is mounted) by oracle/make_golden.py and
oracle/make_constants.py to pin the oracle and to generate tests/golden/*.
Nothing under tests -m gpu, smoke() or bench.py imports this module:
/root/reference does not exist on the GPU box.

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

REFERENCE_ROOT = os.environ.get("EPN_REFERENCE_ROOT", "/root/reference")


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
    """vgtk.cuda.* served by the C oracle (CPU tensors in, CPU tensors out)."""
    from oracle import epn_oracle as O

    pkg = types.ModuleType("vgtk.cuda")
    pkg.__path__ = []
    grouping = types.ModuleType("vgtk.cuda.grouping")
    gathering = types.ModuleType("vgtk.cuda.gathering")
    zpconv = types.ModuleType("vgtk.cuda.zpconv")

    grouping.ball_query = lambda new_xyz, xyz, radius, nsample: O.ball_query(new_xyz, xyz, radius, nsample)
    grouping.furthest_point_sampling = lambda xyz, m: O.furthest_point_sampling(xyz, m)
    gathering.gather_points_forward = lambda pts, idx: O.gather_points_forward(pts, idx)
    gathering.gather_points_backward = lambda g, idx, n: O.gather_points_backward(g, idx, n)
    zpconv.inter_zpconv_forward = O.zp_inter_forward
    zpconv.inter_zpconv_backward = O.zp_inter_backward
    zpconv.intra_zpconv_forward = O.zp_intra_forward
    zpconv.intra_zpconv_backward = O.zp_intra_backward
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


def load_spconvnets():
    """Import SPConvNets.utils.base_so3conv + the three model builders."""
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
