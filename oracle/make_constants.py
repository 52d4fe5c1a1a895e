"""Generate epn_pointcloud_b200/data/so3_constants.npz by running the reference's own
import-time initialisation (vgtk/vgtk/so3conv/functional.py:274-299,
vgtk/vgtk/functional/rotation.py:236-343) through oracle/ref_harness.py.

Run in the build container only:  python -m oracle.make_constants
Outputs: anchors (60,3,3) float32; intra_idx (60,12) int64; kpsphere{24,30,66}
(raw, un-normalised kernel points read from vgtk/vgtk/data/anchors/*.ply).
"""
import os
import numpy as np

from oracle import ref_harness as H


def main():
    vgtk = H.load_reference()
    L = vgtk.so3conv.functional
    root = os.path.join(H.REFERENCE_ROOT, "vgtk", "vgtk", "data", "anchors")
    out = {
        "anchors": np.ascontiguousarray(L.get_anchors(60)).astype(np.float32),
        "intra_idx": np.ascontiguousarray(L.get_intra_idx()).astype(np.int64),
    }
    for n in (24, 30, 66):
        out["kpsphere%d" % n] = H.read_ascii_vertices(os.path.join(root, "kpsphere%d.ply" % n))
    dst = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))),
                       "epn_pointcloud_b200", "data", "so3_constants.npz")
    np.savez(dst, **out)
    print("wrote", dst, {k: v.shape for k, v in out.items()})


if __name__ == "__main__":
    main()
