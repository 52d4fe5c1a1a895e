"""Losses and rotation utilities against outputs of the UNMODIFIED reference
(vgtk/vgtk/loss.py, functional/rotation.py) stored in tests/golden/losses.npz by oracle/make_golden_losses.py.
Host-side torch code: runs on CPU."""
import types

import numpy as np
import pytest
import torch

from conftest import load_golden, rel_err
from epn_pointcloud_b200 import losses as LS


@pytest.fixture(scope="module")
def g():
    return load_golden("losses")


def test_rotation_utilities(g):
    assert rel_err(LS.compute_rotation_matrix_from_quaternion(g["q"]), g["q_R"]) < 1e-6
    assert rel_err(LS.compute_rotation_matrix_from_ortho6d(g["o6"]), g["o6_R"]) < 1e-6
    assert rel_err(LS.so3_mean(g["mean_Rs"], g["mean_w"]), g["mean_R"]) < 1e-5
    assert rel_err(LS.so3_mean(g["mean_Rs"]), g["mean_R_unweighted"]) < 1e-5
    assert rel_err(LS.acos_safe(g["acos_x"]), g["acos_y"]) < 1e-6
    R = LS.compute_rotation_matrix_from_quaternion(torch.randn(5, 4))      # proper rotations
    eye = torch.eye(3).expand(5, 3, 3)
    assert torch.allclose(R @ R.transpose(1, 2), eye, atol=1e-5) and torch.allclose(torch.det(R), torch.ones(5), atol=1e-5)
    assert float(LS.mean_angular_error(R, R).abs().max()) < 2e-2           # acos_safe's linear end: angle(I) ~ 1.4e-2


def _scalars(res, n):
    return np.array([float(v) for v in res[:n]], dtype=np.float64)


@pytest.mark.parametrize("name,loss_type,w,rl", [("cls_default_2d", "default", "cls_w2", "cls_rl1"),
                                                 ("cls_noreg_2d", "no_reg", "cls_w2", "cls_rl1"),
                                                 ("cls_default_3d", "default", "cls_w3", "cls_rl2")])
def test_attention_cross_entropy(g, name, loss_type, w, rl):
    m = LS.AttentionCrossEntropyLoss(loss_type, 0.7)
    got = _scalars(m(g["cls_pred"], g["cls_label"], g[w], g[rl]), 5)
    assert np.allclose(got, g[name].numpy(), rtol=1e-6, atol=1e-7) and m.iter_counter == 1


def test_attention_cross_entropy_schedule(g):
    m = LS.AttentionCrossEntropyLoss("schedule", 0.7)
    m.iter_counter = 500
    got = _scalars(m(g["cls_pred"], g["cls_label"], g["cls_w2"], g["cls_rl1"], pretrain_step=2000), 5)
    assert np.allclose(got, g["cls_schedule_2d"].numpy(), rtol=1e-6, atol=1e-7)
    with pytest.raises(NotImplementedError):
        LS.AttentionCrossEntropyLoss("nope", 1.0)(g["cls_pred"], g["cls_label"], g["cls_w2"], g["cls_rl1"])


def test_multitask_detection_loss(g):
    from epn_pointcloud_b200 import functional as L
    anchors = torch.from_numpy(L.get_anchors(60))
    m = LS.MultiTaskDetectionLoss(anchors, nr=4)
    res = m(g["rot_conf"], g["rot_label"], g["rot_y"], g["rot_gtR"], g["rot_gtT"])            # alignment setting
    assert np.allclose(_scalars(res, 4), g["rot_align_scalars"].numpy(), rtol=2e-5, atol=1e-6)
    assert rel_err(res[4], g["rot_align_err"]) < 1e-4
    res = LS.MultiTaskDetectionLoss(anchors, nr=4)(g["rot_conf1"], g["rot_label1"], g["rot_y1"], g["rot_gtR"])  # canonical
    assert np.allclose(_scalars(res, 4), g["rot_canon_scalars"].numpy(), rtol=2e-5, atol=1e-6)
    assert rel_err(res[4], g["rot_canon_err"]) < 1e-4
    # gradients flow to the regressed residuals and the confidences
    y = g["rot_y"].clone().requires_grad_(True)
    c = g["rot_conf"].clone().requires_grad_(True)
    m(c, g["rot_label"], y, g["rot_gtR"], g["rot_gtT"])[0].backward()
    assert bool(torch.isfinite(y.grad).all()) and float(y.grad.abs().sum()) > 0 and float(c.grad.abs().sum()) > 0


@pytest.mark.parametrize("loss_type", ["soft", "hard", "contrastive"])
def test_triplet_batch_loss(g, loss_type):
    opt = types.SimpleNamespace(device="cpu", train_loss=types.SimpleNamespace(loss_type=loss_type, margin=1.0))
    m = LS.TripletBatchLoss(opt, torch.eye(3)[None])
    got = _scalars(m(g["tri_src"], g["tri_tgt"], None), 4)
    assert np.allclose(got, g["tri_" + loss_type].numpy(), rtol=1e-6, atol=1e-7)
    d = LS.pairwise_distance_matrix(g["tri_src"], g["tri_tgt"])
    assert torch.allclose(d, torch.cdist(g["tri_src"], g["tri_tgt"]), atol=1e-5)


def test_triplet_equivariance_term(g):
    """alpha > 0 (vgtk/vgtk/loss.py:320-428).  Upstream's branch raises for every batch size, so the pins are its own
    single-sample `_interpolate` outputs (golden), the identity-rotation property and the composition of the two
    triplet terms."""
    from epn_pointcloud_b200 import functional as L
    anchors = torch.from_numpy(L.get_anchors(60))
    opt = types.SimpleNamespace(device="cpu", train_loss=types.SimpleNamespace(loss_type="soft", margin=1.0))
    m = LS.TripletBatchLoss(opt, anchors, alpha=0.5)
    # batched interpolation == upstream's per-sample results
    assert rel_err(m._interpolate(g["interp_feat"], g["interp_T"], sigma=0.2), g["interp_out"]) < 1e-5
    # identity rotation with a sharp kernel returns the features themselves
    eye = torch.eye(3)[None].repeat(3, 1, 1)
    assert rel_err(m._interpolate(g["interp_feat"], eye, sigma=1e-3), g["interp_feat"]) < 1e-5
    # total = invariance + alpha * triplet(equi_src, interpolated equi_tgt)
    src, tgt = g["tri_src"][:3], g["tri_tgt"][:3]
    es = g["interp_feat"]
    et = es + 0.05 * torch.randn(es.shape, generator=torch.Generator().manual_seed(3))
    total, inv_info, equi_info = m(src, tgt, eye, es, et)
    inv = LS.TripletBatchLoss(opt, anchors)(src, tgt, None)
    assert abs(float(inv_info[0]) - float(inv[0])) < 1e-7
    d = LS.pairwise_distance_matrix(es.reshape(3, -1), m._interpolate(et, eye, sigma=m.sigma).reshape(3, -1))
    want = torch.nn.functional.softplus(torch.diagonal(d) - LS.batch_hard_negative_mining(d), beta=1.0).mean()
    assert abs(float(equi_info[0]) - float(want)) < 1e-6 and abs(float(total) - float(inv[0]) - 0.5 * float(want)) < 1e-6
    assert float(equi_info[1]) == 1.0   # matching anchors' features retrieve each other
