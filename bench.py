#!/usr/bin/env python
"""Benchmark of the SPConv hot path (BASELINE.json metric):
    point-clouds/sec, ModelNet40 1024-pt 60-anchor SPConv fwd+bwd, 1/2/4/8 B200.

A "step" = one training pass of the classification backbone (7 separable blocks = 7 InterSO3Conv +
7 IntraSO3Conv with their norms / activations / skip branches, BASELINE configs[1]) over one batch
of 32 synthetic clouds per GPU: zero grads, forward, loss, backward, gradient all-reduce (N > 1),
Adam step.

    python bench.py [--gpus N] [--steps K] [--warmup W]              # this framework on N B200s
    python bench.py --impl reference [--steps K] [--warmup W]         # reference CPU path (oracle port)
    torchrun ... bench.py --gpus N ...                                # one rank per GPU (N > 1)

Prints ONE JSON line (rank 0).  See DESIGN.md "Measurement" for every field.
"""
import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

import torch  # noqa: E402

METRIC = "point-clouds/sec, ModelNet40 1024-pt 60-anchor SPConv fwd+bwd"
N_POINTS, N_ANCHORS, KS, KN = 1024, 60, 24, 12
CLASSES = ["index_ops", "inter_group_fwd", "inter_group_bwd_scatter", "intra_group", "channel_gemm", "split_convert",
           "norm_act", "inter_fused_fwd"]


def synthetic_clouds(b, n, seed):
    """randn normalised to the unit sphere surface, centred, max-norm scaled."""
    g = torch.Generator().manual_seed(seed)
    x = torch.randn(b, n, 3, generator=g)
    x = x / x.norm(dim=2, keepdim=True)
    x = x - x.mean(1, keepdim=True)
    return (x / x.norm(dim=2).amax(dim=1).view(b, 1, 1)).contiguous()


def layer_table():
    """(c_in, c_out, p_in, p, k) of the 7 inter layers; each is followed by an intra layer c_out->c_out at p."""
    from epn_pointcloud_b200.blocks import cls_backbone_params
    rows, p_in = [], N_POINTS
    for blk in cls_backbone_params(N_POINTS, N_ANCHORS):
        for l in blk:
            a = l["args"]
            p = -(-p_in // a["stride"])
            rows.append((a["dim_in"], a["dim_out"], p_in, p, a["n_neighbor"]))
            p_in = p
    return rows


def algorithmic_work(batch):
    """Per-step algorithmic flops / bytes per kernel class (formulas of SURVEY.md 8(d), fp32)."""
    A = N_ANCHORS
    gemm_f = group_f = scatter_f = 0.0
    group_b = scatter_b = intra_b = 0.0
    for c_in, c_out, p_in, p, k in layer_table():
        inter_gemm = 2.0 * c_out * c_in * KS * p * A
        intra_gemm = 2.0 * c_out * c_out * KN * p * A
        has_dx = c_in > 1  # layer 0: feats == 1, no dfeats
        gemm_f += inter_gemm * (2 + (1 if has_dx else 0)) + intra_gemm * 3          # fwd + dW (+ dX)
        spatial = 2.0 * c_in * p * A * KS * k + 11.0 * p * A * KS * k
        group_f += spatial                                                          # forward only: dW reads the kept tiles
        scatter_f += spatial if has_dx else 0.0
        grouped = 4.0 * c_in * KS * p * A
        feats_in = 4.0 * c_in * p_in * A if has_dx else 0.0
        group_b += feats_in + 12.0 * p_in + 4.0 * p * k + grouped
        scatter_b += (feats_in + 12.0 * p_in + 4.0 * p * k + grouped) if has_dx else 0.0
        intra_b += 4.0 * c_out * p * A + 4.0 * c_out * KN * p * A                   # training forward gather into kept tiles
    return {"channel_gemm": (gemm_f * batch, None), "inter_group_fwd": (group_f * batch, group_b * batch),
            "inter_group_bwd_scatter": (scatter_f * batch, scatter_b * batch), "intra_group": (None, intra_b * batch)}


class ClockSampler:
    """nvidia-smi clocks + throttle reasons sampled every 200 ms while the timed region runs."""
    Q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.rows, self.proc, self.index = [], None, index

    def __enter__(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q,
                                          "--format=csv,noheader,nounits", "-lms", "200"], stdout=subprocess.PIPE,
                                         stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except OSError:
            self.proc = None
        return self

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def __exit__(self, *a):
        if self.proc is not None:
            self.proc.terminate()
            self.thread.join(timeout=2)

    def summary(self):
        sm = [float(r[0]) for r in self.rows if len(r) >= 6 and r[0].replace(".", "").isdigit()]
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["unavailable"]}
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = [n for i, n in enumerate(names) if any(r[2 + i].lower().startswith("active") for r in self.rows if len(r) >= 6)]
        return {"sm_mhz": statistics.median(sm), "sm_max_mhz": float(self.rows[0][1]), "reasons": reasons,
                "samples": len(sm)}


def synthetic_labels(b, seed):
    return torch.randint(0, 40, (b,), generator=torch.Generator().manual_seed(1000 + seed))


def cpu_port_step(port, x, labels, requires_grad=True):
    """One fwd+bwd of the classification network (backbone + head + cross-entropy) through
    oracle/torch_port.py (the reference's op chain on CPU)."""
    from oracle import torch_port as TP
    layers, hp = port
    for prm, *_ in layers:
        for k in prm:
            prm[k] = prm[k].detach().requires_grad_(requires_grad)
    for k, v in hp.items():
        if k != "anchors":
            hp[k] = [t.detach().requires_grad_(requires_grad) for t in v] if isinstance(v, list) else \
                v.detach().requires_grad_(requires_grad)
    xyz, feats = TP.backbone_forward(x, layers)
    logits, _ = TP.cls_head(xyz, feats, hp)
    loss = torch.nn.functional.cross_entropy(logits, labels)
    loss.backward()
    return float(loss.detach())


def port_of(model):
    from oracle import torch_port as TP
    return TP.layers_from_module(model), TP.head_from_module(model.outblock)


def run_reference(args, rank):
    """--impl reference: the reference's own CPU path (oracle port of its PyTorch op chain; the
    reference is Python and cannot travel to the GPU box) on all host cores."""
    if rank != 0:
        return
    from epn_pointcloud_b200.heads import ClsSO3ConvModel, cls_model_params
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    torch.manual_seed(0)
    model = ClsSO3ConvModel(cls_model_params(N_POINTS, N_ANCHORS))
    port = port_of(model)
    sample = args.ref_batch
    x, labels = synthetic_clouds(sample, N_POINTS, 2), synthetic_labels(sample, 2)
    for _ in range(args.warmup):
        cpu_port_step(port, x, labels)
    t0 = time.perf_counter()
    for _ in range(args.steps):
        cpu_port_step(port, x, labels)
    dt = time.perf_counter() - t0
    value = sample * args.steps / dt
    line = {"impl": "reference", "metric": METRIC, "value": value, "unit": "clouds/s", "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * dt / args.steps,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": "ModelNet40 cls network (7 inter + 7 intra SPConv layers + head) fwd+bwd, 1024 pts, "
                                   "60 anchors (BASELINE configs[1])",
                       "sample": "%d clouds per step (bounded sample of the 32-cloud batch)" % sample},
            "cpu_baseline": {"value": value, "unit": "clouds/s", "cores": cores, "kind": "port",
                             "sample": "%d clouds/step x %d steps, oracle/torch_port.py (reference op chain, torch CPU)" % (sample, args.steps)},
            "e2e": {"value": value, "unit": "clouds/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--batch", type=int, default=32, help="clouds per GPU per step (weak scaling)")
    ap.add_argument("--ref-batch", type=int, default=2, help="clouds per step of the CPU reference arm")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-graph", action="store_true", help="launch every step eagerly instead of replaying a CUDA graph")
    ap.add_argument("--profile-step", action="store_true",
                    help="run warm-up, then ONE step inside cudaProfilerStart/Stop and exit (for ncu)")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "ours" else args.warmup

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        return run_reference(args, rank)

    import torch.distributed as dist
    import epn_pointcloud_b200  # noqa: F401
    from epn_pointcloud_b200 import _lib
    from epn_pointcloud_b200.heads import ClsSO3ConvModel, cls_model_params
    from epn_pointcloud_b200.parallel import FlatGradSync, GraphedTrainStep

    assert torch.cuda.is_available(), "bench.py measures the CUDA path; there is no CPU fallback"
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)
    L = _lib.lib()
    assert L.epn_device_supported() == 1

    torch.manual_seed(0)  # identical weights on every rank
    model = ClsSO3ConvModel(cls_model_params(N_POINTS, N_ANCHORS)).to(dev).train()
    sync = FlatGradSync(model.parameters())
    opt = torch.optim.Adam(model.parameters(), lr=1e-4)
    B = args.batch
    x_host = synthetic_clouds(B, N_POINTS, 2 + rank).pin_memory()
    x_dev = x_host.to(dev)
    x_stage = torch.empty_like(x_dev)
    labels = synthetic_labels(B, 2 + rank).to(dev)

    def eager_step(x):
        sync.zero()
        logits, _ = model(x)
        loss = torch.nn.functional.cross_entropy(logits, labels)
        loss.backward()
        sync.all_reduce_mean()
        opt.step()
        return loss

    step, graphed = eager_step, None
    if not args.no_graph and not args.profile_step:
        # the public training-step helper: forward + loss + backward replayed from ONE CUDA graph
        graphed = GraphedTrainStep(model, lambda out, lab: torch.nn.functional.cross_entropy(out[0], lab), opt, sync,
                                   x_dev, labels, warmup=args.warmup)
        step = lambda x: graphed(x)  # noqa: E731

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(v):
        if world == 1:
            return v
        t = torch.tensor([v], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t)

    for _ in range(args.warmup):
        step(x_dev)
    if args.profile_step:
        # for `ncu --profile-from-start off ...`: exactly one step inside the profiler range, no timing
        torch.cuda.synchronize()
        torch.cuda.profiler.start()
        step(x_dev)
        torch.cuda.synchronize()
        torch.cuda.profiler.stop()
        return
    # ---- device-resident timing ("value")
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    n0 = L.epn_launch_count()
    with ClockSampler(local_rank) as clocks:
        e0.record()
        for _ in range(args.steps):
            loss = step(x_dev)
        e1.record()
        barrier()
    launches = L.epn_launch_count() - n0
    if graphed is not None:  # replayed launches are not seen by the host-side counter: counted once at capture
        launches = graphed.launches_per_replay * args.steps
    ms = max_over_ranks(e0.elapsed_time(e1))
    value = world * B * args.steps / (ms * 1e-3)

    # ---- end to end through the public API with host buffers ("e2e")
    barrier()
    e0.record()
    for _ in range(args.steps):
        x_stage.copy_(x_host, non_blocking=True)          # H2D of the step's input from pinned memory
        loss_host = step(x_stage).item()                  # D2H read of the step's result
    e1.record()
    barrier()
    ms_e2e = max_over_ranks(e0.elapsed_time(e1))
    e2e = world * B * args.steps / (ms_e2e * 1e-3)

    # ---- per-kernel-class device times of ONE extra step (CUDA events at the launch sites, on the
    #      launching stream) -> roofline of the dominant kernel class
    import ctypes
    L.epn_profile_enable(1)
    eager_step(x_dev)
    torch.cuda.synchronize()
    L.epn_profile_enable(0)
    ms_c = (ctypes.c_double * len(CLASSES))()
    n_c = (ctypes.c_longlong * len(CLASSES))()
    L.epn_profile_read(ctypes.cast(ms_c, ctypes.c_void_p), ctypes.cast(n_c, ctypes.c_void_p), len(CLASSES))
    kernel_ms = {c: round(ms_c[i], 3) for i, c in enumerate(CLASSES)}
    kernel_n = {c: int(n_c[i]) for i, c in enumerate(CLASSES)}
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except OSError:
        pass
    hbm_peak = peaks.get("hbm_gbs", 6650.0)
    tf_peak = peaks.get("bf16_tflops_sustained", 1400.0)
    peak_src = "MEASURED_PEAKS.json" if peaks else "fallback"
    work = algorithmic_work(B)
    dom = max(work, key=lambda c: kernel_ms[c])
    flops, nbytes = work[dom]
    t_dom = kernel_ms[dom] * 1e-3
    if dom == "channel_gemm":
        ach = flops / t_dom / 1e12
        roof = {"bound": "tensor", "achieved": ach, "peak": tf_peak, "unit": "TFLOP/s", "frac": ach / tf_peak}
    else:
        ach = nbytes / t_dom / 1e9
        roof = {"bound": "hbm", "achieved": ach, "peak": hbm_peak, "unit": "GB/s", "frac": ach / hbm_peak}
    traffic = None
    try:  # measured DRAM bytes of that kernel class from the committed ncu pass of the same step
        tr = json.load(open(os.path.join(ROOT, "profiles", "traffic_per_step.json")))["classes"][dom]
        traffic = tr["dram_bytes"] / max(tr["launches"], 1)
    except (OSError, KeyError, ValueError):
        pass
    per_launch = max(kernel_n[dom], 1)
    roof["per_launch"] = {"algorithmic": (flops if dom == "channel_gemm" else nbytes) / per_launch,
                          "avg_ms": kernel_ms[dom] / per_launch, "traffic_bytes": traffic}
    roof.update({"traffic": traffic, "kernel": dom, "launches_per_step": kernel_n[dom], "ms_per_step": kernel_ms[dom],
                 "share_of_step": kernel_ms[dom] / (ms / args.steps), "peak_source": peak_src,
                 "note": "algorithmic fp32 flops vs the measured sustained bf16 cuBLAS rate (kernel timed inside a long step)"
                 if dom == "channel_gemm" else "grouping-stage bytes of SURVEY.md 8(d)"})

    line = {"metric": METRIC, "value": value, "unit": "clouds/s", "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": "ModelNet40 cls network (7 inter + 7 intra SPConv layers + head) fwd+bwd+Adam, "
                                   "1024 pts, 60 anchors (BASELINE configs[1])", "clouds_per_gpu": B, "global_batch": B * world,
                       "parallelism": "batch-sharded x%d, one flat-gradient all-reduce" % world,
                       "launch": "eager" if graphed is None else "CUDA graph replay (fwd+loss+bwd), eager all-reduce + Adam",
                       "l2": "no explicit flush: every step streams several GB of activations through the 126 MB L2"},
            "e2e": {"value": e2e, "unit": "clouds/s", "h2d_bytes_per_step": x_host.numel() * 4, "d2h_bytes_per_step": 4,
                    "ms_per_step": ms_e2e / args.steps},
            "gpu_launches": int(launches), "clocks": clocks.summary(), "roofline": roof,
            "kernel_ms_per_step": kernel_ms, "kernel_scopes_per_step": kernel_n, "loss": loss_host}

    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        from oracle import torch_port as TP
        cores = os.cpu_count() or 1
        torch.set_num_threads(cores)
        port = port_of(model)
        xs, ls = synthetic_clouds(args.ref_batch, N_POINTS, 2), synthetic_labels(args.ref_batch, 2)
        t0 = time.perf_counter()
        cpu_port_step(port, xs, ls)
        dt = time.perf_counter() - t0
        line["cpu_baseline"] = {"value": args.ref_batch / dt, "unit": "clouds/s", "cores": cores, "kind": "port",
                                "sample": "%d clouds, 1 fwd+bwd step of the same backbone through oracle/torch_port.py "
                                          "(reference op chain on torch CPU), %.1f s" % (args.ref_batch, dt)}
    if rank == 0:
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
