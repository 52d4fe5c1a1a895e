"""Compile the REFERENCE's own CUDA extensions for sm_100a into oracle/_ref/ (git-ignored binaries).

TEST INFRASTRUCTURE ONLY.  Run in the build container (where /root/reference is mounted):
    python -m oracle.build_ref
The sources are read where they lie (/root/reference/vgtk/vgtk/cuda/*); a temporary patched copy is
made under /tmp because torch >= 2.x rejects `tensor.type()` inside AT_DISPATCH_FLOATING_TYPES (the
22 dispatch sites become `.scalar_type()`) -- nothing from the reference is
written into the repository, only the three compiled modules:
    oracle/_ref/vgtk_ref_grouping.so    ball_query, furthest_point_sampling, anchor_query, ...
    oracle/_ref/vgtk_ref_gathering.so   gather_points_forward/backward
    oracle/_ref/vgtk_ref_zpconv.so      inter/intra_zpconv_forward/backward
They travel to the GPU box with the snapshot and give the `-m gpu` tests a bit-exact GPU-side
pin of the index ops and of the zpconv surface against the reference kernels themselves.
"""
import os
import re
import shutil
import sys
import tempfile

REF_CUDA = os.path.join(os.environ.get("EPN_REFERENCE_ROOT", "/root/reference"), "vgtk", "vgtk", "cuda")
OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "_ref")
EXTS = ("grouping", "gathering", "zpconv")


def module_path(ext):
    return os.path.join(OUT, "vgtk_ref_%s.so" % ext)


def build(force=False):
    if not os.path.isdir(REF_CUDA):
        return False
    from torch.utils.cpp_extension import load
    os.makedirs(OUT, exist_ok=True)
    tmp = tempfile.mkdtemp(prefix="epn_ref_src_")
    try:
        for ext in EXTS:
            if os.path.exists(module_path(ext)) and not force:
                continue
            srcs = []
            for fn in ("%s_cuda.cpp" % ext, "%s_cuda_kernel.cu" % ext):
                text = open(os.path.join(REF_CUDA, fn)).read()
                text = re.sub(r"AT_DISPATCH_FLOATING_TYPES\((\w+)\.type\(\)", r"AT_DISPATCH_FLOATING_TYPES(\1.scalar_type()", text)
                dst = os.path.join(tmp, fn)
                open(dst, "w").write(text)
                srcs.append(dst)
            bdir = os.path.join(tmp, "build_" + ext)
            os.makedirs(bdir)
            name = "vgtk_ref_%s" % ext
            load(name=name, sources=srcs, build_directory=bdir, is_python_module=False, verbose=False,
                 extra_cuda_cflags=["-gencode", "arch=compute_100a,code=sm_100a", "-O3"],
                 extra_cflags=["-O2", "-w"])
            shutil.copy(os.path.join(bdir, name + ".so"), module_path(ext))
    finally:
        shutil.rmtree(tmp, ignore_errors=True)
    return True


def load_ref(ext):
    """Import a prebuilt reference extension (GPU box or here); returns None if absent."""
    path = module_path(ext)
    if not os.path.exists(path):
        return None
    import importlib.util
    import torch  # noqa: F401  (libtorch must be loaded first)
    spec = importlib.util.spec_from_file_location("vgtk_ref_%s" % ext, path)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


if __name__ == "__main__":
    ok = build(force="--force" in sys.argv)
    print("built" if ok else "reference tree not present; nothing built", os.listdir(OUT) if os.path.isdir(OUT) else "")
