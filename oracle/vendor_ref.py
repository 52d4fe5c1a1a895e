"""Make the reference's PYTHON travel to the GPU box: verbatim copy of its two packages into the git-ignored
baseline/_ref/ (the directory the bench contract reserves for the installed reference).

TEST / MEASUREMENT INFRASTRUCTURE ONLY.  Run in the build container (where /root/reference is mounted):
    python -m oracle.vendor_ref
`__graft_entry__.build()` calls it whenever /root/reference is present.  The reference cannot be pip-installed
(its setup.py, vgtk/setup.py:30-34, compiles three CUDA extensions against the torch of 2020 and its imports need
trimesh / plyfile / open3d, none installable here), so the copy stands in for
`pip install --target baseline/_ref /root/reference`.  What is copied: vgtk/vgtk/**/*.py + the anchor / kernel
point .ply data, SPConvNets/**/*.py.  Not copied: the CUDA sources (oracle/build_ref.py compiles those into
oracle/_ref/*.so), media, datasets.  baseline/_ref is listed in .gitignore (never committed) but not in
.gpurunignore, so it ships with the snapshot; bench.py's reference arms import it through oracle/ref_harness.py.
"""
import os
import shutil

SRC = os.environ.get("EPN_REFERENCE_ROOT", "/root/reference")
DST = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "baseline", "_ref")
KEEP = (".py", ".ply", ".json", ".txt")


def vendor(force=False):
    if not os.path.isdir(os.path.join(SRC, "vgtk", "vgtk")):
        return False
    marker = os.path.join(DST, ".vendored_from")
    if os.path.exists(marker) and not force:
        return True
    if os.path.isdir(DST):
        shutil.rmtree(DST)
    for top in (os.path.join("vgtk", "vgtk"), "SPConvNets"):
        for root, dirs, files in os.walk(os.path.join(SRC, top)):
            dirs[:] = [d for d in dirs if d not in ("cuda", "__pycache__")]
            rel = os.path.relpath(root, SRC)
            for fn in files:
                if fn.endswith(KEEP):
                    os.makedirs(os.path.join(DST, rel), exist_ok=True)
                    shutil.copy(os.path.join(root, fn), os.path.join(DST, rel, fn))
    with open(marker, "w") as f:
        f.write("verbatim copy of %s (python + anchor data only); see oracle/vendor_ref.py\n" % SRC)
    return True


if __name__ == "__main__":
    print("vendored" if vendor(force=True) else "reference tree not present; nothing copied", DST)
