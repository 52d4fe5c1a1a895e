"""Multi-GPU plumbing: one process per GPU, clouds sharded across ranks, the gradients of ONE flat fp32 buffer
all-reduced per step (replaces the reference's nn.DataParallel, vgtk/vgtk/app/trainer.py:153-160; BatchNorm
statistics stay per rank = DataParallel replica semantics).  Works with backend "nccl" (NVLink 5 / NVSwitch on the
B200 box) and "gloo" (CPU tests).
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
    """Views every parameter's .grad into one contiguous buffer (cls network: 7.81 M params = 31.3 MB).

    Plain use: `zero()` ... backward ... `all_reduce_mean()` = ONE collective per step.

    `overlap=True` cuts the buffer into buckets of ~`bucket_bytes` (in parameter order: backward produces the last
    layers' gradients first) and launches each bucket's all-reduce on a side stream as soon as autograd has
    accumulated its last gradient (post-accumulate hooks), so the collectives of the head and the deep blocks run
    under the backward kernels of the shallow ones; `all_reduce_mean()` then only waits for the side stream.
    The fork/join is plain stream-event ordering, so it can be captured into a CUDA graph together with the backward
    pass (GraphedTrainStep does).  On CPU tensors (gloo tests) the bucket collectives run synchronously in the hook.
    """

    def __init__(self, params, process_group=None, overlap=False, bucket_bytes=8 << 20):
        self.params = [p for p in params if p.requires_grad]
        self.group = process_group
        total = sum(p.numel() for p in self.params)
        ref = self.params[0]
        self.flat = torch.zeros(total, dtype=ref.dtype, device=ref.device)
        off = 0
        self._spans = []
        for p in self.params:
            n = p.numel()
            p.grad = self.flat[off:off + n].view_as(p)
            self._spans.append((off, off + n))
            off += n
        self._views = [p.grad for p in self.params]
        self.overlap = False
        if overlap:
            self._setup_overlap(bucket_bytes)

    # ------------------------------------------------------------------ bucketed, overlapped reduction
    def _distributed(self):
        return dist.is_available() and dist.is_initialized() and dist.get_world_size(self.group) > 1

    def _setup_overlap(self, bucket_bytes):
        self.overlap = True
        per = max(int(bucket_bytes) // self.flat.element_size(), 1)
        self._buckets, self._bucket_of = [], []   # [lo, hi, n_params], parameter -> bucket
        lo = cnt = 0
        for i, (a, b) in enumerate(self._spans):
            self._bucket_of.append(len(self._buckets))
            cnt += 1
            if b - lo >= per or i == len(self._spans) - 1:
                self._buckets.append([lo, b, cnt])
                lo, cnt = b, 0
        self._pending = [b[2] for b in self._buckets]
        self._side = torch.cuda.Stream(self.flat.device) if self.flat.is_cuda else None
        self._launched = False
        for i, p in enumerate(self.params):
            p.register_post_accumulate_grad_hook(self._make_hook(i))

    def _make_hook(self, i):
        def hook(p):
            if not self._distributed():
                return
            g, v = p.grad, self._views[i]
            if g is not None and g.data_ptr() != v.data_ptr():   # autograd bound a private gradient: bring it home
                v.copy_(g)
                p.grad = v
            k = self._bucket_of[i]
            self._pending[k] -= 1
            if self._pending[k] == 0:
                self._reduce_bucket(k)
        return hook

    def _reduce_bucket(self, k):
        lo, hi, _ = self._buckets[k]
        chunk = self.flat[lo:hi]
        ws = dist.get_world_size(self.group)
        self._launched = True
        if self._side is None:
            dist.all_reduce(chunk, op=dist.ReduceOp.SUM, group=self.group)
            chunk.mul_(1.0 / ws)
            return
        self._side.wait_stream(torch.cuda.current_stream(self.flat.device))
        with torch.cuda.stream(self._side):
            dist.all_reduce(chunk, op=dist.ReduceOp.AVG, group=self.group)

    # ------------------------------------------------------------------ public
    def zero(self):
        self.flat.zero_()
        if self.overlap:
            self._pending = [b[2] for b in self._buckets]
            self._launched = False

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
        if not self._distributed():
            return
        ws = dist.get_world_size(self.group)
        if self.overlap and local_units is None:
            # buckets whose gradients all arrived are already reduced (or in flight on the side stream); a bucket
            # with a parameter that received no gradient this step is reduced here
            for k, left in enumerate(self._pending):
                if left > 0:
                    self._reduce_bucket(k)
            if self._side is not None:
                torch.cuda.current_stream(self.flat.device).wait_stream(self._side)
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


def _tree_map(fn, obj):
    if isinstance(obj, torch.Tensor):
        return fn(obj)
    if isinstance(obj, (tuple, list)):
        return type(obj)(_tree_map(fn, o) for o in obj)
    return obj


def _tree_copy(dst, src):
    if isinstance(dst, torch.Tensor):
        if src is not None and src.data_ptr() != dst.data_ptr():
            dst.copy_(src, non_blocking=True)
    elif isinstance(dst, (tuple, list)):
        for d, s in zip(dst, src):
            _tree_copy(d, s)


class GraphedTrainStep:
    """One training step (zero grads, forward, loss, backward, and -- with an overlapping FlatGradSync -- the
    bucketed gradient all-reduce) captured ONCE into a CUDA graph and replayed.

    The SPConv stack issues several hundred small-to-medium kernel launches per step; replayed from a graph the
    GPU never waits for the host, which matters whenever the caller synchronises every step (reads the loss, feeds
    fresh host data) and when the per-GPU batch is small (strong scaling).  The optimizer step runs eagerly after
    the replay, so any torch optimizer works unchanged.

        step = GraphedTrainStep(model, loss_fn, optimizer, sync, x_example, labels_example)
        loss = step(x, labels)            # x / labels: CUDA or pinned-host tensors of the captured shapes

    `loss_fn(model_output, labels) -> scalar tensor`; `labels` is a tensor or a tuple of tensors; `sync` is a
    FlatGradSync (gradients live in its flat buffer, zeroed inside the graph).  Shapes, dtypes and the module's
    training mode are frozen at capture.
    """

    def __init__(self, model, loss_fn, optimizer, sync, x_example, labels_example, warmup=3):
        self.model, self.loss_fn, self.optimizer, self.sync = model, loss_fn, optimizer, sync
        self.x = x_example.detach().clone()
        self.labels = _tree_map(lambda t: t.detach().clone(), labels_example)
        # the collectives are captured with the backward pass only when they were launched from its hooks
        self.reduce_in_graph = bool(getattr(sync, "overlap", False)) and sync._distributed() and sync.flat.is_cuda
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
        if self.reduce_in_graph:
            self.sync.all_reduce_mean()   # joins the side stream the bucket collectives were forked onto
        return loss.detach()

    def __call__(self, x=None, labels=None):
        if x is not None and x.data_ptr() != self.x.data_ptr():
            self.x.copy_(x, non_blocking=True)
        if labels is not None:
            _tree_copy(self.labels, labels)
        self.graph.replay()
        if not self.reduce_in_graph:
            self.sync.all_reduce_mean()
        self.optimizer.step()
        return self.loss
