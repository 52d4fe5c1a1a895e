#!/usr/bin/env python
"""Benchmark of the SPConv hot path (BASELINE.json metric):
    point-clouds/sec, ModelNet40 1024-pt 60-anchor SPConv fwd+bwd, 1/2/4/8 B200.

A "step" = one training pass of a shipped network over one batch of synthetic clouds: zero grads, forward, loss,
backward, gradient all-reduce (N > 1), Adam step.  The headline line is BASELINE configs[1] (classification
network, 32 clouds per GPU, weak scaling); the same run also measures, and reports inside the same JSON line,
  * `forward`        inference forward of the same network (the ">= 10x the reference GPU forward" target),
  * `strong_scaling` the same network with the GLOBAL batch fixed at 32 clouds (SURVEY 8e partitioning: 32/N per GPU),
  * `other_configs`  BASELINE configs[2] (rotation network, 32 pairs = 64 clouds, strong scaling) and configs[3]
                     (3DMatch descriptor network, 16 patches of 2048 points, strong scaling),
  * `reference_gpu`  the REFERENCE's own modules + its own CUDA extensions on the same GPU (N = 1 only),
  * `rooflines`      achieved fraction of the measured peaks for every kernel class of the library.

    python bench.py [--gpus N] [--steps K] [--warmup W]                      # this framework on N B200s
    python bench.py --config reg|inv ...                                      # make another network the headline
    python bench.py --impl reference [--steps K] [--warmup W]                 # reference CPU path
    torchrun ... bench.py --gpus N ...                                        # one rank per GPU (N > 1)

Prints ONE JSON line (rank 0).  See DESIGN.md "Measurement" for every field.
"""
import argparse
import json
import math
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
DTYPE = "f32 (tensor-core operands split hi/lo: bf16x3 in training, fp16x3 in the no_grad forward; fp32 accumulate; SIMT stages fp32)"


# ---------------------------------------------------------------------------------------------- synthetic data
def synthetic_clouds(b, n, seed):
    """randn normalised to the unit sphere surface, centred, max-norm scaled."""
    g = torch.Generator().manual_seed(seed)
    x = torch.randn(b, n, 3, generator=g)
    x = x / x.norm(dim=2, keepdim=True)
    x = x - x.mean(1, keepdim=True)
    return (x / x.norm(dim=2).amax(dim=1).view(b, 1, 1)).contiguous()


def synthetic_labels(b, seed):
    return torch.randint(0, 40, (b,), generator=torch.Generator().manual_seed(1000 + seed))


def random_rotations(b, seed):
    g = torch.Generator().manual_seed(seed)
    q = torch.randn(b, 4, generator=g)
    q = q / q.norm(dim=1, keepdim=True)
    w, x, y, z = q.unbind(1)
    return torch.stack([1 - 2 * (y * y + z * z), 2 * (x * y - z * w), 2 * (x * z + y * w),
                        2 * (x * y + z * w), 1 - 2 * (x * x + z * z), 2 * (y * z - x * w),
                        2 * (x * z - y * w), 2 * (y * z + x * w), 1 - 2 * (x * x + y * y)], 1).view(b, 3, 3)


def synthetic_pairs(n_pairs, n, seed):
    """(source, target) pairs, source = target rotated by a random rotation T, with the
    labels the reference's loader derives from T (SPConvNets/datasets/modelnet40.py:131-152,
    vgtk/vgtk/functional/rotation.py:521-526): per-anchor target label and residual rotation."""
    from epn_pointcloud_b200 import functional as L
    tgt = synthetic_clouds(n_pairs, n, seed)
    T = random_rotations(n_pairs, seed + 1)
    src = torch.einsum("bij,bnj->bni", T, tgt)
    anchors = torch.from_numpy(L.get_anchors(60))
    t_from = torch.einsum("abc,nbj,ijk->naick", anchors, T, anchors)            # [n, a, i, 3, 3]
    label = torch.einsum("naikk->nai", t_from).argmax(2)                         # [n, 60]
    R = torch.gather(t_from, 2, label.view(n_pairs, 60, 1, 1, 1).expand(-1, -1, 1, 3, 3)).squeeze(2)
    return torch.stack([src, tgt], 1).contiguous(), (label.contiguous(), R.contiguous(), T.contiguous())


def synthetic_patches(n_pairs, n, seed):
    """points uniform in a ball of radius 0.4 (search_radius); matching patch = the same
    points under a random rotation.  Returns [2*n_pairs, n, 3]: sources first, then their targets."""
    g = torch.Generator().manual_seed(seed)
    d = torch.randn(n_pairs, n, 3, generator=g)
    src = 0.4 * d / d.norm(dim=2, keepdim=True) * torch.rand(n_pairs, n, 1, generator=g) ** (1.0 / 3.0)
    tgt = torch.einsum("bij,bnj->bni", random_rotations(n_pairs, seed + 1), src)
    return torch.cat([src, tgt], 0).contiguous()


# ---------------------------------------------------------------------------------------------- workloads
class Workload:
    """One BASELINE config: network, synthetic batch, loss.  `units` = what the metric counts (clouds / patches);
    `items` = what is sharded across ranks (clouds, pairs, patch pairs)."""

    def __init__(self, name):
        from epn_pointcloud_b200 import heads, losses
        self.name = name
        if name == "cls":
            self.n_points, self.global_items, self.units_per_item = 1024, 32, 1
            self.params = heads.cls_model_params(1024, 60)
            self.model_cls = heads.ClsSO3ConvModel
            self.desc = ("ModelNet40 cls network (7 inter + 7 intra SPConv layers + head) fwd+bwd+Adam, 1024 pts, "
                         "60 anchors (BASELINE configs[1])")
            self.loss = lambda out, lab: torch.nn.functional.cross_entropy(out[0], lab)
        elif name == "reg":
            self.n_points, self.global_items, self.units_per_item = 1024, 32, 2
            self.params = heads.reg_model_params(1024, 60)
            self.model_cls = heads.RegSO3ConvModel
            self.desc = ("ModelNet40 relative-rotation network (7+7 SPConv layers, K=64/32, pair head) fwd+bwd+Adam, "
                         "32 pairs = 64 clouds of 1024 pts, 60 anchors (BASELINE configs[2])")
            self._metric = None   # built on the device in build()
            self.loss = lambda out, lab: self._metric(out[0], lab[0], out[1], lab[1], lab[2])[0]
        elif name == "inv":
            self.n_points, self.global_items, self.units_per_item = 2048, 8, 2
            self.params = heads.inv_model_params(2048, 60)
            self.model_cls = heads.InvSO3ConvModel
            self.desc = ("3DMatch descriptor network (8+8 SPConv layers, K=128 first layer, 64-d descriptor) "
                         "fwd+bwd+Adam, 16 patches (8 src/tgt pairs) of 2048 pts, 60 anchors (BASELINE configs[3])")

            def triplet(out, lab):
                d = out[0]
                n = d.shape[0] // 2
                dist_ = losses.pairwise_distance_matrix(d[:n], d[n:])
                # batch-hard mining without boolean indexing (no host sync: the step is captured into a CUDA graph)
                neg = (dist_ + torch.eye(n, device=d.device, dtype=d.dtype) * 1e9).min(1)[0]
                return torch.nn.functional.softplus(torch.diagonal(dist_) - neg, beta=1.0).mean()   # TripletBatchLoss, 'soft'
            self.loss = triplet
        else:
            raise ValueError(name)

    def build(self, dev):
        torch.manual_seed(0)  # identical weights on every rank
        m = self.model_cls(self.params).to(dev).train()
        if self.name == "reg":
            from epn_pointcloud_b200 import functional as L, losses
            self._metric = losses.MultiTaskDetectionLoss(torch.from_numpy(L.get_anchors(60)).to(dev), nr=4)
            self._metric.with_error = False   # the logging-only angular error needs an SVD (host sync): not part of the step
        return m

    def batch(self, n_items, seed):
        """-> (x [host], labels [host tensor / tuple / None-placeholder tensor])"""
        if self.name == "cls":
            return synthetic_clouds(n_items, self.n_points, seed), synthetic_labels(n_items, seed)
        if self.name == "reg":
            return synthetic_pairs(n_items, self.n_points, seed)
        return synthetic_patches(n_items, self.n_points, seed), torch.zeros(1)

    def layer_table(self):
        """(c_in, c_out, p_in, p, k) of the inter layers; each is followed by an intra layer c_out->c_out at p."""
        rows, p_in = [], self.n_points
        for blk in self.params["backbone"]:
            for l in blk:
                a = l["args"]
                p = -(-p_in // a["stride"])
                rows.append((a["dim_in"], a["dim_out"], p_in, p, a["n_neighbor"]))
                p_in = p
        return rows

    def algorithmic_work(self, clouds, fused=True, fused_bwd=False):
        """Per-step algorithmic flops / bytes per kernel class (formulas of SURVEY.md 8(d), fp32) + the fused-layer
        totals of the forward."""
        A = N_ANCHORS
        gemm_f = group_f = scatter_f = 0.0
        group_b = scatter_b = intra_b = norm_b = split_b = 0.0
        fused_b = fused_f = 0.0
        ifu_b = ifu_f = 0.0
        for c_in, c_out, p_in, p, k in self.layer_table():
            inter_gemm = 2.0 * c_out * c_in * KS * p * A
            intra_gemm = 2.0 * c_out * c_out * KN * p * A
            has_dx = c_in > 1  # layer 0: feats == 1, no dfeats
            # channel-GEMM kernels: dW (+ dX) of the inter conv, its forward only for layer 0 (the other layers'
            # forward GEMM runs inside the fused inter kernel and is counted there), fwd + dX + dW of the intra conv
            gemm_f += inter_gemm * 2 + intra_gemm * 3
            spatial = 2.0 * c_in * p * A * KS * k + 11.0 * p * A * KS * k
            scatter_f += spatial if has_dx else 0.0
            if fused_bwd and has_dx and k <= (16 if fused_bwd < 2 else 32) and c_out % 64 == 0 and c_out <= 256 and c_in % 4 == 0 and p % 2 == 0:
                gemm_f -= inter_gemm      # the data-gradient GEMM of these layers runs inside the fused backward kernel
                scatter_f += inter_gemm
            grouped = 4.0 * c_in * KS * p * A
            feats_in = 4.0 * c_in * p_in * A if has_dx else 0.0
            if not (fused and has_dx):   # grouping-only kernels: every layer without the fused kernel, else layer 0 only
                group_f += spatial                                                      # forward only: dW reads the kept tiles
                group_b += feats_in + 12.0 * p_in + 4.0 * p * k + grouped
            scatter_b += (feats_in + 12.0 * p_in + 4.0 * p * k + grouped) if has_dx else 0.0
            intra_b += 4.0 * c_out * p * A + 4.0 * c_out * KN * p * A                   # training forward gather into kept tiles
            # pass-level bytes of the HBM-bound helper kernels (what each pass must read + write; under the fused-layer
            # accounting of SURVEY 8(d) all of this is overhead with zero algorithmic bytes)
            t_out, t_skip = 4.0 * c_out * p * A, 4.0 * c_in * p * A
            norm_b += 25.0 * t_out    # 3 norms per block: fwd stats + apply (+ residual) = 10 passes, bwd reduce + apply = 15
            split_b += 2.0 * (6.0 * t_out + 2.0 * t_skip)   # dout of 3 convs in 2 orientations; skip-conv input fwd + dW
            fused_b += feats_in + 12.0 * p_in + 4.0 * c_out * p * A + 4.0 * p * k + 8.0 * c_out * p * A
            fused_f += spatial + inter_gemm + intra_gemm
            if has_dx:   # the layers the fused inter-conv kernel runs (layer 0 has its own single-channel kernel)
                ifu_b += feats_in + 12.0 * p_in + 4.0 * c_out * p * A + 4.0 * p * k      # fused InterSO3Conv bytes
                ifu_f += spatial + inter_gemm
        return {"channel_gemm": (gemm_f * clouds, None), "inter_group_fwd": (group_f * clouds, group_b * clouds),
                "inter_group_bwd_scatter": (scatter_f * clouds, scatter_b * clouds), "intra_group": (None, intra_b * clouds),
                "fused_forward": (fused_f * clouds, fused_b * clouds), "inter_fused_fwd": (ifu_f * clouds, ifu_b * clouds),
                "norm_act": (None, norm_b * clouds), "split_convert": (None, split_b * clouds)}


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


# ---------------------------------------------------------------------------------------------- reference arms
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


def reference_cls_model(device, gpu_ops):
    """The REFERENCE's own classification network (SPConvNets/models/cls_so3net_pn.py, unmodified Python from
    baseline/_ref or /root/reference, through oracle/ref_harness.py), default-initialised under seed 0.  Its three
    native ops come from the reference's own CUDA extensions (oracle/_ref, gpu_ops=True) or from the C oracle."""
    from oracle import ref_harness
    if not ref_harness.available():
        return None
    ref_harness.load_spconvnets(gpu_ops=gpu_ops)
    torch.manual_seed(0)
    return ref_harness.build_cls_model(N_POINTS, N_ANCHORS).to(device).train()


def run_reference(args, rank):
    """--impl reference: the reference's own CPU implementation of the path on all host cores -- its unmodified
    modules (kind "reference") when the vendored tree baseline/_ref (or /root/reference) is present, else the oracle
    port of its op chain (kind "port")."""
    if rank != 0:
        return
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    sample = args.ref_batch
    x, labels = synthetic_clouds(sample, N_POINTS, 2), synthetic_labels(sample, 2)
    model = reference_cls_model("cpu", gpu_ops=False)
    if model is not None:
        kind, how = "reference", "the reference's own modules (ClsSO3ConvModel, unmodified Python) on torch CPU"

        def step():
            model.zero_grad(set_to_none=True)
            logits, _ = model(x)
            loss = torch.nn.functional.cross_entropy(logits, labels)
            loss.backward()
            return float(loss)
    else:
        from epn_pointcloud_b200.heads import ClsSO3ConvModel, cls_model_params
        kind, how = "port", "oracle/torch_port.py (reference op chain, torch CPU)"
        torch.manual_seed(0)
        port = port_of(ClsSO3ConvModel(cls_model_params(N_POINTS, N_ANCHORS)))

        def step():
            return cpu_port_step(port, x, labels)
    for _ in range(args.warmup):
        step()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        step()
    dt = time.perf_counter() - t0
    value = sample * args.steps / dt
    line = {"impl": "reference", "metric": METRIC, "value": value, "unit": "clouds/s", "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * dt / args.steps,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": Workload("cls").desc.replace("+Adam", ""),
                       "sample": "%d clouds per step (bounded sample of the 32-cloud batch)" % sample},
            "cpu_baseline": {"value": value, "unit": "clouds/s", "cores": cores, "kind": kind,
                             "sample": "%d clouds/step x %d steps, %s" % (sample, args.steps, how)},
            "e2e": {"value": value, "unit": "clouds/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line), flush=True)


# ---------------------------------------------------------------------------------------------- measurement
class Runner:
    def __init__(self, args):
        import torch.distributed as dist
        self.dist = dist
        self.args = args
        self.rank = int(os.environ.get("RANK", "0"))
        self.world = int(os.environ.get("WORLD_SIZE", "1"))
        self.local_rank = int(os.environ.get("LOCAL_RANK", "0"))
        torch.cuda.set_device(self.local_rank)
        self.dev = torch.device("cuda", self.local_rank)
        if self.world > 1:
            os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
            dist.init_process_group("nccl", device_id=self.dev)

    def barrier(self):
        if self.world > 1:
            self.dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(self, v):
        if self.world == 1:
            return v
        t = torch.tensor([v], dtype=torch.float64, device=self.dev)
        self.dist.all_reduce(t, op=self.dist.ReduceOp.MAX)
        return float(t)

    def timed(self, fn, steps):
        self.barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(steps):
            out = fn()
        e1.record()
        self.barrier()
        return self.max_over_ranks(e0.elapsed_time(e1)), out

    def train_setup(self, wl, items_per_rank, graph=True, warmup=3):
        """-> dict(step, eager_step, x_host, x_dev, x_stage, labels, graphed, model, sync)"""
        from epn_pointcloud_b200.parallel import FlatGradSync, GraphedTrainStep
        dev = self.dev
        model = wl.build(dev)
        sync = FlatGradSync(model.parameters(), overlap=self.world > 1)
        opt = torch.optim.Adam(model.parameters(), lr=1e-4)
        x_host, labels = wl.batch(items_per_rank, 2 + self.rank)
        x_host = x_host.pin_memory()
        x_dev = x_host.to(dev)
        labels = tuple(t.to(dev) for t in labels) if isinstance(labels, tuple) else labels.to(dev)

        def eager_step(x):
            sync.zero()
            loss = wl.loss(model(x), labels)
            loss.backward()
            sync.all_reduce_mean()
            opt.step()
            return loss

        step, graphed = eager_step, None
        if graph:
            try:
                graphed = GraphedTrainStep(model, wl.loss, opt, sync, x_dev, labels, warmup=warmup)
                step = lambda x: graphed(x)  # noqa: E731
            except Exception as e:  # a loss with a host sync (SVD of the rotation loss) cannot be captured: run eagerly
                torch.cuda.synchronize()
                self.graph_fallback = repr(e)[:160]
                graphed = None
        for _ in range(warmup):
            step(x_dev)
        return {"step": step, "eager_step": eager_step, "x_host": x_host, "x_dev": x_dev,
                "x_stage": torch.empty_like(x_dev), "graphed": graphed, "model": model, "sync": sync, "opt": opt}

    def measure_training(self, wl, items_per_rank, steps, warmup, graph=True, e2e=True):
        L = self.L
        S = self.train_setup(wl, items_per_rank, graph, warmup)
        step, x_dev = S["step"], S["x_dev"]
        units = self.world * items_per_rank * wl.units_per_item
        n0 = L.epn_launch_count()
        with ClockSampler(self.local_rank) as clocks:
            ms, _ = self.timed(lambda: step(x_dev), steps)
        launches = L.epn_launch_count() - n0
        if S["graphed"] is not None:  # replayed launches are not seen by the host-side counter: counted once at capture
            launches = S["graphed"].launches_per_replay * steps
        res = {"value": units * steps / (ms * 1e-3), "ms_per_step": ms / steps, "launches": int(launches),
               "launch": "CUDA graph replay" if S["graphed"] is not None else "eager",
               "clocks": clocks.summary(), "units_per_gpu": items_per_rank * wl.units_per_item, "setup": S}
        if e2e:
            x_host, x_stage = S["x_host"], S["x_stage"]

            def e2e_step():
                x_stage.copy_(x_host, non_blocking=True)          # H2D of the step's input from pinned memory
                return step(x_stage).item()                       # D2H read of the step's result
            ms2, loss_host = self.timed(e2e_step, steps)
            res["e2e"] = {"value": units * steps / (ms2 * 1e-3), "unit": "clouds/s",
                          "h2d_bytes_per_step": x_host.numel() * 4, "d2h_bytes_per_step": 4, "ms_per_step": ms2 / steps}
            res["loss"] = loss_host
        return res

    def measure_forward(self, model, x_dev, steps, warmup, eval_mode=False):
        """no_grad forward.  eval_mode False: module in training mode (BatchNorm batch statistics) -- the mode the
        reference's forward is timed in (reference_gpu_block); True: evaluation mode (running statistics)."""
        model.train(not eval_mode)

        def fwd():
            with torch.no_grad():
                return model(x_dev)
        for _ in range(warmup):
            fwd()
        ms, _ = self.timed(fwd, steps)
        model.train()
        return ms / steps


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--config", default="cls", choices=["cls", "reg", "inv"], help="network of the headline line")
    ap.add_argument("--scaling", default=None, choices=["weak", "strong"],
                    help="cls default weak (32 clouds per GPU); reg / inv default strong (BASELINE global batch)")
    ap.add_argument("--batch", type=int, default=None, help="items (clouds / pairs / patch pairs) per GPU (weak) or in total (strong)")
    ap.add_argument("--ref-batch", type=int, default=2, help="clouds per step of the CPU reference arm")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-extras", action="store_true", help="headline measurement only (no forward / strong / other configs / reference_gpu)")
    ap.add_argument("--no-graph", action="store_true", help="launch every step eagerly instead of replaying a CUDA graph")
    ap.add_argument("--profile-step", action="store_true",
                    help="run warm-up, then ONE step inside cudaProfilerStart/Stop and exit (for ncu)")
    ap.add_argument("--profile-forward", action="store_true", help="like --profile-step for one inference forward")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "ours" else args.warmup

    rank = int(os.environ.get("RANK", "0"))
    if args.impl == "reference":
        return run_reference(args, rank)

    import epn_pointcloud_b200  # noqa: F401
    from epn_pointcloud_b200 import _lib
    assert torch.cuda.is_available(), "bench.py measures the CUDA path; there is no CPU fallback"
    R = Runner(args)
    L = R.L = _lib.lib()
    assert L.epn_device_supported() == 1
    world = R.world

    wl = Workload(args.config)
    scaling = args.scaling or ("weak" if args.config == "cls" else "strong")
    total = args.batch if args.batch is not None else wl.global_items
    items = total if scaling == "weak" else max(total // world, 1)
    graph = not args.no_graph and not args.profile_step and not args.profile_forward

    if args.profile_step or args.profile_forward:
        S = R.train_setup(wl, items, graph=False, warmup=args.warmup)
        torch.cuda.synchronize()
        torch.cuda.profiler.start()
        if args.profile_forward:   # same mode as the `forward` measurement: no_grad, module in training mode
            with torch.no_grad():
                S["model"](S["x_dev"])
        else:
            S["eager_step"](S["x_dev"])
        torch.cuda.synchronize()
        torch.cuda.profiler.stop()
        return

    main_res = R.measure_training(wl, items, args.steps, args.warmup, graph=graph)
    S = main_res.pop("setup")
    ms_step = main_res["ms_per_step"]

    # ---- per-kernel-class device times of ONE extra eager step (CUDA events at the launch sites, on the
    #      launching stream) -> roofline of every kernel class
    import ctypes

    def class_times(fn):
        L.epn_profile_enable(1)
        fn()
        torch.cuda.synchronize()
        L.epn_profile_enable(0)
        ms_c = (ctypes.c_double * len(CLASSES))()
        n_c = (ctypes.c_longlong * len(CLASSES))()
        L.epn_profile_read(ctypes.cast(ms_c, ctypes.c_void_p), ctypes.cast(n_c, ctypes.c_void_p), len(CLASSES))
        return {c: round(ms_c[i], 3) for i, c in enumerate(CLASSES)}, {c: int(n_c[i]) for i, c in enumerate(CLASSES)}

    kernel_ms, kernel_n = class_times(lambda: S["eager_step"](S["x_dev"]))
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except OSError:
        pass
    hbm_peak = peaks.get("hbm_gbs", 6650.0)
    tf_peak = peaks.get("bf16_tflops_sustained", 1400.0)
    peak_src = "MEASURED_PEAKS.json" if peaks else "fallback"
    clouds_per_gpu = items * wl.units_per_item
    work = wl.algorithmic_work(clouds_per_gpu, fused=kernel_ms.get("inter_fused_fwd", 0.0) > 0.0,
                               fused_bwd=int(L.epn_get_fused_inter_bwd()))
    traffic_tab = {}
    try:  # measured DRAM bytes per class from the committed ncu pass of the same step
        traffic_tab = json.load(open(os.path.join(ROOT, "profiles", "traffic_per_step.json")))
    except (OSError, ValueError):
        pass

    def roof_of(cls):
        flops, nbytes = work[cls]
        t = kernel_ms.get(cls, 0.0) * 1e-3
        if t <= 0:
            return None
        per_launch = max(kernel_n[cls], 1)
        if cls in ("channel_gemm", "inter_fused_fwd"):
            ach = flops / t / 1e12
            r = {"bound": "tensor", "achieved": ach, "peak": tf_peak, "unit": "TFLOP/s", "frac": ach / tf_peak,
                 "note": "algorithmic fp32 flops vs the measured sustained bf16 cuBLAS rate; the bf16x3 scheme caps frac at 1/3"}
            if cls == "inter_fused_fwd":   # both accountings of SURVEY 8(d); the tensor one binds (larger ideal time)
                r["hbm_frac"] = nbytes / t / 1e9 / hbm_peak
                r["algorithmic_bytes"] = nbytes
                r["note"] = ("fused InterSO3Conv (gather + spatial contraction + channel GEMM): algorithmic fp32 flops vs the "
                             "sustained bf16 rate (x3 MMAs per product: cap 1/3); HBM accounting in hbm_frac (not binding)")
            alg = flops
        else:
            ach = nbytes / t / 1e9
            r = {"bound": "hbm", "achieved": ach, "peak": hbm_peak, "unit": "GB/s", "frac": ach / hbm_peak,
                 "note": "grouping-stage bytes of SURVEY.md 8(d)" if cls not in ("norm_act", "split_convert") else
                         "pass-level bytes (what each pass must read + write): how close the passes run to HBM speed; under the "
                         "fused-layer accounting of SURVEY 8(d) these passes carry no algorithmic bytes at all"}
            alg = nbytes
        traffic = None
        try:
            tr = traffic_tab["classes"][cls]
            traffic = tr["dram_bytes"] / max(tr["launches"], 1)
        except (KeyError, TypeError):
            pass
        r.update({"traffic": traffic, "traffic_source": traffic_tab.get("source") if traffic is not None else None,
                  "kernel": cls, "launches_per_step": kernel_n[cls], "ms_per_step": kernel_ms[cls],
                  "share_of_step": kernel_ms[cls] / ms_step, "peak_source": peak_src,
                  "per_launch": {"algorithmic": alg / per_launch, "avg_ms": kernel_ms[cls] / per_launch}})
        return r

    rooflines = {c: roof_of(c) for c in ("channel_gemm", "inter_fused_fwd", "inter_group_fwd", "inter_group_bwd_scatter", "intra_group",
                                         "norm_act", "split_convert")}
    rooflines = {c: r for c, r in rooflines.items() if r is not None}
    dom = max(rooflines, key=lambda c: rooflines[c]["ms_per_step"])

    line = {"metric": METRIC if args.config == "cls" else "clouds/sec, " + wl.desc, "value": main_res["value"],
            "unit": "clouds/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_step,
            "higher_is_better": True, "scaling": scaling, "vs_baseline": None, "dtype": DTYPE, "data": "synthetic",
            "config": {"workload": wl.desc, "clouds_per_gpu": clouds_per_gpu, "global_batch": clouds_per_gpu * world,
                       "parallelism": "batch-sharded x%d, bucketed flat-gradient all-reduce overlapped with backward" % world,
                       "launch": "eager" if not graph else "CUDA graph replay (fwd+loss+bwd%s), eager Adam"
                                 % (" + bucketed all-reduce" if world > 1 else ""),
                       "l2": "no explicit flush: every step streams several GB of activations through the 126 MB L2"},
            "e2e": main_res["e2e"], "gpu_launches": main_res["launches"], "clocks": main_res["clocks"],
            "roofline": rooflines[dom], "rooflines": rooflines,
            "kernel_ms_per_step": kernel_ms, "kernel_scopes_per_step": kernel_n, "loss": main_res["loss"]}

    extras = not args.no_extras
    if extras:
        # ---- inference forward of the same network, same batch (the >= 10x reference-GPU-forward target)
        fwd_ms = R.measure_forward(S["model"], S["x_dev"], max(args.steps // 2, 3), 2)
        fwd_eval_ms = R.measure_forward(S["model"], S["x_dev"], max(args.steps // 2, 3), 2, eval_mode=True)
        fwd_kernel_ms, _ = class_times(lambda: R.measure_forward(S["model"], S["x_dev"], 1, 0))
        ff, fb = work["fused_forward"]
        line["forward"] = {"value": world * clouds_per_gpu / (fwd_ms * 1e-3), "unit": "clouds/s", "ms": fwd_ms,
                           "mode": "no_grad, module in training mode (BatchNorm batch statistics) -- as the reference forward is timed",
                           "eval_mode_ms": fwd_eval_ms, "eval_mode_clouds_per_s": world * clouds_per_gpu / (fwd_eval_ms * 1e-3),
                           "kernel_ms": {k: v for k, v in fwd_kernel_ms.items() if v > 0},
                           "vs_fused_layer_roofline": {
                               "algorithmic_GB": fb / 1e9, "algorithmic_TFLOP": ff / 1e12,
                               "hbm_frac": fb / (fwd_ms * 1e-3) / 1e9 / hbm_peak,
                               "bf16x3_tensor_frac": 3.0 * ff / (fwd_ms * 1e-3) / 1e12 / tf_peak,
                               "note": "whole forward against the fused-layer accounting of SURVEY 8(d): bytes = feats in + "
                                       "out of every conv, flops = spatial + channel GEMMs (x3 for the bf16x3 MMA scheme)"}}
    del S
    torch.cuda.empty_cache()

    if extras and args.config == "cls":
        # ---- strong scaling of the headline network: global batch 32 (32/N clouds per GPU)
        if world > 1:
            r = R.measure_training(wl, max(32 // world, 1), args.steps, args.warmup, graph=graph, e2e=False)
            r.pop("setup")
            line["strong_scaling"] = {"value": r["value"], "unit": "clouds/s", "ms_per_step": r["ms_per_step"],
                                      "global_batch": 32, "clouds_per_gpu": r["units_per_gpu"]}
            torch.cuda.empty_cache()
        # ---- the other BASELINE configs (strong scaling: BASELINE's global batch sharded over the N GPUs)
        other = {}
        for name in ("reg", "inv"):
            w2 = Workload(name)
            it = max(w2.global_items // world, 1)
            try:
                r = R.measure_training(w2, it, max(args.steps // 2, 3), 3, graph=graph, e2e=False)
                S2 = r.pop("setup")
                f_ms = R.measure_forward(S2["model"], S2["x_dev"], 3, 1)
                other[name] = {"workload": w2.desc, "value": r["value"], "unit": "clouds/s", "ms_per_step": r["ms_per_step"],
                               "scaling": "strong", "clouds_per_gpu": r["units_per_gpu"], "global_batch": r["units_per_gpu"] * world,
                               "gpu_launches_per_step": r["launches"] // max(args.steps // 2, 3), "launch": r["launch"],
                               "forward_clouds_per_s": world * r["units_per_gpu"] / (f_ms * 1e-3), "forward_ms": f_ms}
                del S2, r
            except Exception as e:  # a failure here must not lose the headline line
                other[name] = {"error": repr(e)[:300]}
            torch.cuda.empty_cache()
        line["other_configs"] = other

    if extras and rank == 0 and world == 1 and args.config == "cls":
        # ---- the reference's GPU path on this GPU: its own modules + its own CUDA extensions (oracle/_ref)
        try:
            line["reference_gpu"] = reference_gpu_block(R, clouds_per_gpu)
            if "forward" in line and "forward_clouds_per_s" in line["reference_gpu"]:
                line["forward"]["vs_reference_gpu_forward"] = line["forward"]["value"] / line["reference_gpu"]["forward_clouds_per_s"]
                line["vs_reference_gpu_fwd_bwd"] = line["value"] / line["reference_gpu"]["fwd_bwd_clouds_per_s"]
        except Exception as e:
            line["reference_gpu"] = {"unavailable": repr(e)[:300]}
        torch.cuda.empty_cache()

    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        line["cpu_baseline"] = cpu_baseline_block(args)
    if rank == 0:
        print(json.dumps(line), flush=True)
    if world > 1:
        R.dist.destroy_process_group()


def reference_gpu_block(R, batch):
    """Reference modules + reference CUDA kernels on the GPU: forward (no_grad) at the engine's batch, fwd+bwd at
    the largest batch of (batch, 12 = the authors' setting, run_modelnet.py:10) that fits."""
    model = reference_cls_model(R.dev, gpu_ops=True)
    kind = "reference modules (unmodified Python) + reference CUDA extensions (oracle/_ref)"
    if model is None:
        raise RuntimeError("baseline/_ref (vendored reference Python) not present")
    x = synthetic_clouds(batch, N_POINTS, 2).to(R.dev)
    labels = synthetic_labels(batch, 2).to(R.dev)
    out = {"kind": kind, "batch": batch}

    def fwd():
        with torch.no_grad():
            return model(x)
    fwd()
    ms, _ = R.timed(fwd, 3)
    out["forward_ms"] = ms / 3
    out["forward_clouds_per_s"] = batch / (ms / 3 * 1e-3)
    # parity on the bench's own batch: this engine with the reference model's weights against the reference's output
    # (same check as tests/test_gpu_parity2.py::test_cls_network_b32_vs_reference_modules_on_gpu; bar 1e-4)
    try:
        from epn_pointcloud_b200.heads import ClsSO3ConvModel, cls_model_params
        ours = ClsSO3ConvModel(cls_model_params(N_POINTS, N_ANCHORS)).to(R.dev).train()
        ours.load_state_dict(model.state_dict(), strict=True)
        tf32 = torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32
        torch.backends.cudnn.allow_tf32 = torch.backends.cuda.matmul.allow_tf32 = False   # the reference side in true fp32
        try:
            with torch.no_grad():
                lo, fo = ours(x)
                lr, fr = model(x)
        finally:
            torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32 = tf32

        def rel(a, b):
            return float((a.double() - b.double()).abs().max() / b.double().abs().max())
        out["parity"] = {"what": "no_grad forward of this engine (reference weights) vs the reference modules + kernels, same %d clouds" % batch,
                         "head_feature_max_rel_err": rel(fo, fr), "logits_max_rel_err": rel(lo, lr), "bar": 1e-4}
        del ours, lo, fo, lr, fr
    except Exception as e:  # parity is reported, never allowed to lose the timing block
        out["parity"] = {"error": repr(e)[:200]}
    torch.cuda.empty_cache()
    for b2 in (batch, 12, 8, 4):
        if b2 > batch:
            continue
        try:
            xb, lb = x[:b2], labels[:b2]

            def fb():
                model.zero_grad(set_to_none=True)
                torch.nn.functional.cross_entropy(model(xb)[0], lb).backward()
            fb()
            ms, _ = R.timed(fb, 3)
            out["fwd_bwd_batch"] = b2
            out["fwd_bwd_ms"] = ms / 3
            out["fwd_bwd_clouds_per_s"] = b2 / (ms / 3 * 1e-3)
            break
        except torch.OutOfMemoryError:
            model.zero_grad(set_to_none=True)
            torch.cuda.empty_cache()
    return out


def cpu_baseline_block(args):
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    xs, ls = synthetic_clouds(args.ref_batch, N_POINTS, 2), synthetic_labels(args.ref_batch, 2)
    model = None
    try:
        model = reference_cls_model("cpu", gpu_ops=False)
    except Exception:
        model = None
    t0 = time.perf_counter()
    if model is not None:
        kind, how = "reference", "the reference's own modules (unmodified Python via oracle/ref_harness.py) on torch CPU"
        torch.nn.functional.cross_entropy(model(xs)[0], ls).backward()
    else:
        from epn_pointcloud_b200.heads import ClsSO3ConvModel, cls_model_params
        kind, how = "port", "oracle/torch_port.py (reference op chain on torch CPU)"
        torch.manual_seed(0)
        cpu_port_step(port_of(ClsSO3ConvModel(cls_model_params(N_POINTS, N_ANCHORS))), xs, ls)
    dt = time.perf_counter() - t0
    return {"value": args.ref_batch / dt, "unit": "clouds/s", "cores": cores, "kind": kind,
            "sample": "%d clouds, 1 fwd+bwd step of the cls network through %s, %.1f s" % (args.ref_batch, how, dt)}


if __name__ == "__main__":
    main()
