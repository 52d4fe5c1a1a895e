"""Golden vectors for the losses / rotation utilities, produced by the UNMODIFIED
reference (vgtk/vgtk/loss.py, vgtk/vgtk/functional/rotation.py) on CPU through oracle/ref_harness.py.
Build-container only:  python -m oracle.make_golden_losses   -> tests/golden/losses.npz

TEST INFRASTRUCTURE ONLY.  The reference's quaternion / 6-D helpers move a constant to the GPU with `.cuda()`
(rotation.py:385,449); for this CPU run `torch.Tensor.cuda` is patched to the identity, nothing else is touched.
"""
import types

import numpy as np
import torch

from oracle import ref_harness as H
from oracle.make_golden import npy, save


def main():
    vgtk = H.load_reference()
    import vgtk.loss as RL
    import vgtk.functional as RF
    import vgtk.so3conv.functional as SF
    torch.Tensor.cuda = lambda self, *a, **k: self   # see module docstring
    g = torch.Generator().manual_seed(41)
    out = {}

    # rotation utilities
    q = torch.randn(7, 4, generator=g)
    o6 = torch.randn(7, 6, generator=g)
    out["q"], out["q_R"] = npy(q), npy(RF.compute_rotation_matrix_from_quaternion(q))
    out["o6"], out["o6_R"] = npy(o6), npy(RF.compute_rotation_matrix_from_ortho6d(o6))
    Rs = RF.compute_rotation_matrix_from_quaternion(torch.randn(3 * 5, 4, generator=g)).view(3, 5, 3, 3)
    wts = torch.rand(3, 5, generator=g)
    out["mean_Rs"], out["mean_w"], out["mean_R"], out["mean_R_unweighted"] = npy(Rs), npy(wts), npy(RF.so3_mean(Rs, wts)), npy(RF.so3_mean(Rs))
    x = torch.linspace(-1.2, 1.2, 49)
    from vgtk.spconv.functional import acos_safe
    out["acos_x"], out["acos_y"] = npy(x), npy(acos_safe(x))

    # classification loss
    pred = torch.randn(6, 40, generator=g)
    label = torch.randint(0, 40, (6,), generator=g)
    w2 = torch.randn(6, 60, generator=g)
    w3 = torch.randn(6, 8, 60, generator=g)
    rl1 = torch.randint(0, 60, (6,), generator=g)
    rl2 = torch.randint(0, 60, (6, 60), generator=g)
    out.update(cls_pred=npy(pred), cls_label=npy(label), cls_w2=npy(w2), cls_w3=npy(w3), cls_rl1=npy(rl1), cls_rl2=npy(rl2))
    for name, lt, w, rl in (("cls_default_2d", "default", w2, rl1), ("cls_noreg_2d", "no_reg", w2, rl1), ("cls_default_3d", "default", w3, rl2)):
        m = RL.AttentionCrossEntropyLoss(lt, 0.7)
        out[name] = np.array([float(v) for v in m(pred, label, w, rl)], dtype=np.float64)
    m = RL.AttentionCrossEntropyLoss("schedule", 0.7)
    m.iter_counter = 500
    out["cls_schedule_2d"] = np.array([float(v) for v in m(pred, label, w2, rl1, pretrain_step=2000)], dtype=np.float64)

    # rotation loss, alignment setting (the shipped RegSO3ConvModel) and canonical setting
    anchors = torch.from_numpy(SF.get_anchors(60))
    b = 3
    conf = torch.softmax(torch.randn(b, 60, 60, generator=g), dim=1)
    y = torch.randn(b, 4, 60, 60, generator=g)
    lab = torch.randint(0, 60, (b, 60), generator=g)
    gtR = RF.compute_rotation_matrix_from_quaternion(torch.randn(b * 60, 4, generator=g)).view(b, 60, 3, 3)
    gtT = RF.compute_rotation_matrix_from_quaternion(torch.randn(b, 4, generator=g))
    m = RL.MultiTaskDetectionLoss(anchors, nr=4)
    res = m(conf, lab, y, gtR, gtT)
    out.update(rot_conf=npy(conf), rot_y=npy(y), rot_label=npy(lab), rot_gtR=npy(gtR), rot_gtT=npy(gtT),
               rot_align_scalars=np.array([float(v) for v in res[:4]], dtype=np.float64), rot_align_err=npy(res[4]))
    conf1 = torch.softmax(torch.randn(b, 60, generator=g), dim=1)
    y1 = torch.randn(b, 4, 60, generator=g)
    lab1 = torch.randint(0, 60, (b,), generator=g)
    res = RL.MultiTaskDetectionLoss(anchors, nr=4)(conf1, lab1, y1, gtR)
    out.update(rot_conf1=npy(conf1), rot_y1=npy(y1), rot_label1=npy(lab1),
               rot_canon_scalars=np.array([float(v) for v in res[:4]], dtype=np.float64), rot_canon_err=npy(res[4]))

    # triplet loss
    src = torch.nn.functional.normalize(torch.randn(9, 16, generator=g), dim=1)
    tgt = torch.nn.functional.normalize(src + 0.3 * torch.randn(9, 16, generator=g), dim=1)
    out.update(tri_src=npy(src), tri_tgt=npy(tgt))
    for lt in ("soft", "hard", "contrastive"):
        opt = types.SimpleNamespace(device="cpu", train_loss=types.SimpleNamespace(loss_type=lt, margin=1.0))
        res = RL.TripletBatchLoss(opt, anchors)(src, tgt, None)
        out["tri_" + lt] = np.array([float(v) for v in res], dtype=np.float64)
    # the rotated-feature interpolation of the equivariance term: upstream's _interpolate only runs for ONE sample
    # (its flattened index is reshaped as if the batch were 1, vgtk/vgtk/loss.py:408-409), so three single-sample calls
    opt = types.SimpleNamespace(device="cpu", train_loss=types.SimpleNamespace(loss_type="soft", margin=1.0))
    tl = RL.TripletBatchLoss(opt, anchors, alpha=0.5)
    feat = torch.randn(3, 8, 60, generator=g)
    Ts = RF.compute_rotation_matrix_from_quaternion(torch.randn(3, 4, generator=g))
    outs = [tl._interpolate(feat[i:i + 1], Ts[i:i + 1], sigma=0.2) for i in range(3)]
    out.update(interp_feat=npy(feat), interp_T=npy(Ts), interp_out=npy(torch.cat(outs, 0)))
    save("losses", **out)


if __name__ == "__main__":
    main()
