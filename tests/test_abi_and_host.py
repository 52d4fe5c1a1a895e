"""CPU: the C-ABI library builds, loads and exports every symbol include/epn_b200.h declares;
host-side logic (sharding, flat-gradient all-reduce over gloo, error behaviour without a GPU)."""
import ctypes
import os
import re
import socket
import subprocess
import sys

import pytest
import torch

from conftest import ROOT


def _declared_symbols():
    text = open(os.path.join(ROOT, "include", "epn_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(epn_[a-z0-9_]+)\s*\(", text)))


def test_library_builds_and_exports_every_declared_symbol():
    from epn_pointcloud_b200 import _lib
    so = _lib.build()
    assert os.path.exists(so)
    lib = ctypes.CDLL(so)
    declared = _declared_symbols()
    assert len(declared) >= 25
    for name in declared:
        assert hasattr(lib, name), "header declares %s but the library does not export it" % name
    assert sorted(_lib.SIGNATURES) == declared, "python SIGNATURES table out of sync with the header"
    exported = subprocess.check_output(["nm", "-D", "--defined-only", so]).decode()
    extra = [l.split()[-1] for l in exported.splitlines() if " T " in l and not l.split()[-1].startswith("epn_")]
    assert extra == [] or all(e.startswith("_") for e in extra), extra


def test_version_and_argument_errors_without_gpu():
    from epn_pointcloud_b200 import _lib
    L = _lib.lib()
    assert L.epn_version() == 101
    # NULL pointers and bad extents are rejected before anything touches a device
    rc = L.epn_ball_query_f32(None, None, None, 1, 8, 8, 0.1, 4, None)
    assert rc == -1 and b"NULL" in L.epn_last_error()
    buf = (ctypes.c_float * 64)()
    p = ctypes.cast(buf, ctypes.c_void_p)
    rc = L.epn_ball_query_f32(p, p, p, 0, 8, 8, 0.1, 4, None)
    assert rc == -2 and b"must be > 0" in L.epn_last_error()
    rc = L.epn_inter_so3conv_fwd_f32(None, p, p, p, p, p, 0.1, p, p, p, 16, None, 0, None, 1, 4, 4, 8, 8, 4, 60, 24, None)
    assert rc == -1  # feats NULL with c_in != 1
    # kept operand tiles: one 128-row tile per 128 grouped columns, 4 bytes (bf16 hi+lo) per element, and
    # only for shapes the tile kernels cover with whole tiles
    assert L.epn_inter_so3conv_grouped_bytes(2, 4, 64, 16, 60, 24) == 2 * 64 * 60 * 96 * 4
    assert L.epn_inter_so3conv_grouped_bytes(2, 4, 64, 16, 20, 24) == 0  # 20 anchors: generic path
    assert L.epn_inter_so3conv_grouped_bytes(1, 4, 10, 16, 60, 24) == 0  # 600 columns: not whole tiles
    assert L.epn_intra_so3conv_grouped_bytes(2, 8, 64, 60, 12) == 2 * 64 * 60 * 96 * 4
    wsb = L.epn_inter_so3conv_workspace_bytes(2, 4, 8, 64, 64, 16, 60, 24, 0)
    slab = 2 * 4 * 24 * 64 * 60 * 4  # one fp32 slab + its bf16 hi/lo operand tiles + weight tiles
    assert wsb % 256 == 0 and 2 * slab <= wsb <= 4 * slab
    assert L.epn_get_gemm_backend() in (0, 1)
    assert L.epn_fps_workspace_bytes(4, 1024) == 0 and L.epn_fps_workspace_bytes(4, 20000) == 4 * 20000 * 4


def test_product_path_refuses_cpu_tensors():
    from epn_pointcloud_b200 import ops
    with pytest.raises(RuntimeError, match="CUDA tensor"):
        ops.ball_query(torch.zeros(1, 3, 8), torch.zeros(1, 3, 8), 0.1, 4)
    with pytest.raises(RuntimeError, match="CUDA tensor"):
        ops.intra_so3conv_fwd(torch.zeros(1, 2, 4, 60), torch.zeros(60, 12, dtype=torch.int32), torch.zeros(3, 24))


def test_product_package_never_imports_the_oracle():
    for fn in os.listdir(os.path.join(ROOT, "epn_pointcloud_b200")):
        if fn.endswith(".py"):
            src = open(os.path.join(ROOT, "epn_pointcloud_b200", fn)).read()
            assert "import oracle" not in src and "from oracle" not in src, fn


def test_shard_ranges():
    from epn_pointcloud_b200.parallel import shard_pairs, shard_range
    for n, ws in ((32, 8), (32, 3), (5, 8), (16, 2)):
        spans = [shard_range(n, r, ws) for r in range(ws)]
        assert spans[0][0] == 0 and spans[-1][1] == n
        assert all(spans[i][1] == spans[i + 1][0] for i in range(ws - 1))
        sizes = [hi - lo for lo, hi in spans]
        assert max(sizes) - min(sizes) <= 1
    assert shard_pairs(32, 1, 8) == (8, 16)


_WORKER = r"""
import os, sys, torch, torch.distributed as dist
sys.path.insert(0, sys.argv[1])
from epn_pointcloud_b200.parallel import FlatGradSync, shard_range
dist.init_process_group("gloo", init_method="tcp://127.0.0.1:%s" % sys.argv[2], rank=int(sys.argv[3]), world_size=2)
torch.manual_seed(0)
model = torch.nn.Sequential(torch.nn.Linear(6, 5), torch.nn.Linear(5, 3))
sync = FlatGradSync(model.parameters())
x = torch.arange(8 * 6, dtype=torch.float32).view(8, 6) / 10
lo, hi = shard_range(8, dist.get_rank(), 2)
sync.zero()
model(x[lo:hi]).pow(2).sum().backward()
sync.all_reduce_mean()
ref = torch.nn.Sequential(torch.nn.Linear(6, 5), torch.nn.Linear(5, 3))
ref.load_state_dict(model.state_dict())
(ref(x).pow(2).sum() / 2).backward()
for p, q in zip(model.parameters(), ref.parameters()):
    assert torch.allclose(p.grad, q.grad, rtol=1e-5, atol=1e-6), (p.grad, q.grad)
    assert p.grad.data_ptr() >= sync.flat.data_ptr()
# optimizer.zero_grad() (set_to_none=True) drops the flat views: the next all-reduce must still see every gradient
model.zero_grad(set_to_none=True)
assert all(p.grad is None for p in model.parameters())
model(x[lo:hi]).pow(2).sum().backward()           # autograd allocates private gradients
sync.all_reduce_mean()
for p, q in zip(model.parameters(), ref.parameters()):
    assert torch.allclose(p.grad, q.grad, rtol=1e-5, atol=1e-6)
    assert sync.flat.data_ptr() <= p.grad.data_ptr() < sync.flat.data_ptr() + sync.flat.numel() * 4
# unequal shards (7 clouds over 2 ranks = 4 + 3): per-rank MEAN losses, weighted by the shard sizes
lo, hi = shard_range(7, dist.get_rank(), 2)
sync.rebind(); sync.zero()
model(x[lo:hi]).pow(2).sum(1).mean().backward()
sync.all_reduce_mean(local_units=hi - lo)
ref.zero_grad()
ref(x[:7]).pow(2).sum(1).mean().backward()
for p, q in zip(model.parameters(), ref.parameters()):
    assert torch.allclose(p.grad, q.grad, rtol=1e-5, atol=1e-6), (p.grad, q.grad)
# bucketed reduction launched from the post-accumulate hooks (overlap=True; synchronous on CPU tensors):
# 4 parameters in 3 buckets, one of them left without a gradient in the second step
model2 = torch.nn.Sequential(torch.nn.Linear(6, 5), torch.nn.Linear(5, 3))
model2.load_state_dict(ref.state_dict())
sync2 = FlatGradSync(model2.parameters(), overlap=True, bucket_bytes=64)
assert len(sync2._buckets) >= 2
lo, hi = shard_range(8, dist.get_rank(), 2)
for frozen in (False, True):
    sync2.zero()
    h = model2[0](x[lo:hi])
    out = model2[1](h.detach() if frozen else h)   # frozen: layer 0 receives no gradient this step
    out.pow(2).sum().backward()
    sync2.all_reduce_mean()
    ref.zero_grad()
    hr = ref[0](x)
    (ref[1](hr.detach() if frozen else hr).pow(2).sum() / 2).backward()
    for p, q in zip(model2.parameters(), ref.parameters()):
        want = torch.zeros_like(p) if q.grad is None else q.grad
        assert torch.allclose(p.grad, want, rtol=1e-5, atol=1e-6), (frozen, p.grad, want)
dist.destroy_process_group()
print("ok")
"""


def test_flat_grad_all_reduce_gloo_world2(tmp_path):
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    script = tmp_path / "worker.py"
    script.write_text(_WORKER)
    procs = [subprocess.Popen([sys.executable, str(script), ROOT, str(port), str(r)], stdout=subprocess.PIPE,
                              stderr=subprocess.STDOUT) for r in range(2)]
    outs = [p.communicate(timeout=180)[0].decode() for p in procs]
    for p, o in zip(procs, outs):
        assert p.returncode == 0 and "ok" in o, o


def test_checkpoint_io_in_the_reference_format(tmp_path):
    """f4: a reference checkpoint (plain CPU state_dict, optionally with DataParallel's `module.` prefix or a
    {"model": ...} wrapper, vgtk/vgtk/app/trainer.py:177-223) loads key for key into the mirrored model, a mismatch is
    reported, and save_checkpoint writes what the reference's _save_network would."""
    import torch
    from conftest import load_golden
    from epn_pointcloud_b200.blocks import SO3ConvBackbone
    from epn_pointcloud_b200.checkpoint import load_checkpoint, save_checkpoint
    g = load_golden("backbone_small")
    ref_sd = g.state_dict()                       # produced by the reference's own modules (oracle/make_golden.py)
    path = str(tmp_path / "ref_net_10.pth")
    torch.save(ref_sd, path)
    model = SO3ConvBackbone(g["params"], 60)
    assert load_checkpoint(model, path) == ([], [])
    for k, v in ref_sd.items():
        assert torch.equal(model.state_dict()[k].reshape(v.shape), v), k
    # DataParallel-style keys inside a {"model": ...} wrapper
    model2 = SO3ConvBackbone(g["params"], 60)
    load_checkpoint(model2, {"model": {"module." + k: v for k, v in ref_sd.items()}, "epoch": 3})
    assert all(torch.equal(model2.state_dict()[k].reshape(v.shape), v) for k, v in ref_sd.items())
    # a checkpoint of another architecture is refused with the offending keys named
    bad = dict(ref_sd)
    k0 = next(k for k in bad if k.endswith("basic_conv.W"))
    bad[k0] = bad[k0][:, :-1]
    bad["extra.weight"] = torch.zeros(1)
    try:
        load_checkpoint(model2, bad)
        assert False, "mismatch not reported"
    except RuntimeError as e:
        assert "extra.weight" in str(e) and k0 in str(e)
    # round trip: what save_checkpoint writes is a plain CPU state_dict with exactly the reference's keys
    out = save_checkpoint(torch.nn.DataParallel(model) if False else model, str(tmp_path / "out.pth"))
    back = torch.load(out, weights_only=True)
    assert list(back.keys()) == list(ref_sd.keys()) and all(not v.is_cuda for v in back.values())
    assert all(torch.equal(back[k].reshape(ref_sd[k].shape), ref_sd[k]) for k in ref_sd)


def test_fused_backward_knob_levels_and_bench_accounting():
    """The data-gradient knob is a clamped level (0 / 1 / 2) readable back without a GPU; bench.py's per-class
    accounting moves the data-gradient GEMM flops of the layers the fused kernel covers from the GEMM class to the
    backward-data class without creating or losing any."""
    from epn_pointcloud_b200 import _lib
    L = _lib.lib()
    old = L.epn_get_fused_inter_bwd()
    try:
        for given, want in ((0, 0), (1, 1), (2, 2), (7, 2), (-3, 0)):
            L.epn_set_fused_inter_bwd(given)
            assert L.epn_get_fused_inter_bwd() == want
    finally:
        L.epn_set_fused_inter_bwd(old)
    sys.path.insert(0, ROOT)
    import bench
    wl = bench.Workload("cls")
    tot = []
    for level in (0, 1, 2):
        w = wl.algorithmic_work(32, fused=True, fused_bwd=level)
        tot.append((w["channel_gemm"][0], w["inter_group_bwd_scatter"][0]))
    assert tot[0][0] > tot[1][0] > tot[2][0]                      # more layers leave the GEMM class with every level
    for g, s_ in tot[1:]:
        assert abs((g + s_) - (tot[0][0] + tot[0][1])) < 1e-6 * (tot[0][0] + tot[0][1])
