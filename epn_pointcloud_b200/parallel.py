"""Multi-GPU plumbing: one process per GPU, clouds sharded across ranks, ONE all-reduce of a
flat fp32 gradient buffer per step (replaces the reference's nn.DataParallel,
vgtk/vgtk/app/trainer.py:153-160; BatchNorm statistics stay per rank = DataParallel replica
semantics).  Works with backend "nccl" (NVLink 5 / NVSwitch on the B200 box) and "gloo" (CPU tests).
"""
import torch
import torch.distributed as dist


def shard_range(n_units, rank, world_size):
    """Contiguous, balanced [lo, hi) of `n_units` independent clouds for `rank`."""
    base, rem = divmod(n_units, world_size)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def shard_pairs(n_pairs, rank, world_size):
    """Rotation-estimation batches hold (source, target) pairs: a pair is never split
    (SPConvNets/models/reg_so3net.py:33).  Returns the cloud range [2*lo, 2*hi)."""
    lo, hi = shard_range(n_pairs, rank, world_size)
    return 2 * lo, 2 * hi


class FlatGradSync:
    """Views every parameter's .grad into one contiguous buffer so that a step needs a single
    collective (cls backbone: 7.67 M params = 30.7 MB)."""

    def __init__(self, params, process_group=None):
        self.params = [p for p in params if p.requires_grad]
        self.group = process_group
        total = sum(p.numel() for p in self.params)
        ref = self.params[0]
        self.flat = torch.zeros(total, dtype=ref.dtype, device=ref.device)
        off = 0
        for p in self.params:
            n = p.numel()
            p.grad = self.flat[off:off + n].view_as(p)
            off += n

        self._views = [p.grad for p in self.params]

    def zero(self):
        self.flat.zero_()

    def rebind(self):
        """Point every parameter's .grad back at its slice of the flat buffer.  `optimizer.zero_grad()` (default
        set_to_none=True) or `model.zero_grad()` drop the views: the next backward would then allocate private
        gradients that neither the all-reduce nor a graph replay writing into the flat buffer would see.  A private
        gradient that already holds data is copied in, so nothing is lost."""
        for p, v in zip(self.params, self._views):
            g = p.grad
            if g is None:
                p.grad = v
            elif g.data_ptr() != v.data_ptr():
                v.copy_(g)
                p.grad = v

    def all_reduce_mean(self, local_units=None):
        """Sum over ranks then divide; no-op for a single process.  With `local_units` (clouds this rank's loss
        was averaged over) the result is the gradient of the mean over ALL clouds even when the shards are not
        equal (shard_range with a remainder); without it the shards are assumed equal (mean of per-rank means)."""
        self.rebind()
        if not (dist.is_available() and dist.is_initialized()):
            return
        ws = dist.get_world_size(self.group)
        if ws == 1:
            return
        if local_units is None:
            dist.all_reduce(self.flat, op=dist.ReduceOp.SUM, group=self.group)
            self.flat.mul_(1.0 / ws)
            return
        n = torch.tensor([float(local_units)], dtype=self.flat.dtype, device=self.flat.device)
        self.flat.mul_(n)
        dist.all_reduce(n, op=dist.ReduceOp.SUM, group=self.group)
        dist.all_reduce(self.flat, op=dist.ReduceOp.SUM, group=self.group)
        self.flat.div_(n)


class GraphedTrainStep:
    """One training step (zero grads, forward, loss, backward) captured ONCE into a CUDA graph and replayed.

    The SPConv stack issues ~3000 small-to-medium kernel launches per step; replayed from a graph the GPU never
    waits for the host, which matters whenever the caller synchronises every step (reads the loss, feeds
    fresh host data).  The gradient all-reduce (if a process group is up) and the optimizer step run eagerly
    after the replay, so any torch optimizer works unchanged.

        step = GraphedTrainStep(model, loss_fn, optimizer, sync, x_example, labels_example)
        loss = step(x, labels)            # x / labels: CUDA or pinned-host tensors of the captured shapes

    `loss_fn(model_output, labels) -> scalar tensor`; `sync` is a FlatGradSync (gradients live in its flat
    buffer, zeroed inside the graph).  Shapes, dtypes and the module's training mode are frozen at capture.
    """

    def __init__(self, model, loss_fn, optimizer, sync, x_example, labels_example, warmup=3):
        self.model, self.loss_fn, self.optimizer, self.sync = model, loss_fn, optimizer, sync
        self.x = x_example.detach().clone()
        self.labels = labels_example.detach().clone()
        cur = torch.cuda.current_stream()
        side = torch.cuda.Stream()
        side.wait_stream(cur)
        # the warm-up passes and the capture pass run the model in its current mode on the example batch: module
        # buffers they would advance (BatchNorm running statistics, num_batches_tracked) and the RNG state are
        # snapshotted and restored, so constructing the step leaves the model as it found it
        buffers = [(b, b.detach().clone()) for b in model.buffers()]
        rng = torch.cuda.get_rng_state(self.x.device)
        with torch.cuda.stream(side):   # warm-up off the default stream: lazy initialisations, allocator pools
            for _ in range(max(int(warmup), 1)):
                self._fwd_bwd()
        cur.wait_stream(side)
        torch.cuda.synchronize()
        from . import _lib
        n0 = _lib.lib().epn_launch_count()
        self.graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(self.graph):
            self.loss = self._fwd_bwd()
        self.launches_per_replay = int(_lib.lib().epn_launch_count() - n0)  # this library's kernels inside the graph
        with torch.no_grad():
            for b, saved in buffers:
                b.copy_(saved)
        torch.cuda.set_rng_state(rng, self.x.device)
        sync.zero()

    def _fwd_bwd(self):
        self.sync.rebind()
        self.sync.zero()
        loss = self.loss_fn(self.model(self.x), self.labels)
        loss.backward()
        return loss.detach()

    def __call__(self, x=None, labels=None):
        if x is not None and x.data_ptr() != self.x.data_ptr():
            self.x.copy_(x, non_blocking=True)
        if labels is not None and labels.data_ptr() != self.labels.data_ptr():
            self.labels.copy_(labels, non_blocking=True)
        self.graph.replay()
        self.sync.all_reduce_mean()
        self.optimizer.step()
        return self.loss
