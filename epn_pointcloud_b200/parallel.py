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

    def zero(self):
        self.flat.zero_()

    def all_reduce_mean(self):
        """Sum over ranks then divide by the world size; no-op for a single process."""
        if not (dist.is_available() and dist.is_initialized()):
            return
        ws = dist.get_world_size(self.group)
        if ws == 1:
            return
        for p in self.params:  # a backward pass may have re-bound .grad: copy stragglers back
            if p.grad is not None and p.grad.data_ptr() < self.flat.data_ptr() or \
               p.grad is not None and p.grad.data_ptr() >= self.flat.data_ptr() + self.flat.numel() * self.flat.element_size():
                raise RuntimeError("parameter gradient left the flat buffer")
        dist.all_reduce(self.flat, op=dist.ReduceOp.SUM, group=self.group)
        self.flat.mul_(1.0 / ws)


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

    def _fwd_bwd(self):
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
