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
