"""Checkpoint I/O in the reference's format (the part of it that touches the path).

The reference trainer saves `model.state_dict()` moved to the CPU with `torch.save` and resumes with
`torch.load` + `load_state_dict` (vgtk/vgtk/app/trainer.py:188-223); the authors' published `.pth` files are such
state_dicts.  The modules of this package keep the reference's sub-module names, parameter / buffer names and shapes,
so those files load key for key; the helpers below only add what a user switching over needs around that:
`module.` prefixes of checkpoints written from an `nn.DataParallel` wrapper, `{"model": state_dict, ...}` wrappers (the
commented-out variant at trainer.py:177-186), strict checking with a readable report, and a save that writes exactly
what the reference would have written (so the reference can resume from a checkpoint trained here).
"""
import collections

import torch


def _unwrap(obj):
    if isinstance(obj, dict) and "model" in obj and isinstance(obj["model"], dict):
        obj = obj["model"]          # {"model": ..., "optimizer": ..., "epoch": ...}
    if isinstance(obj, dict) and "state_dict" in obj and isinstance(obj["state_dict"], dict):
        obj = obj["state_dict"]
    if not isinstance(obj, dict):
        raise TypeError("checkpoint does not hold a state_dict (got %s)" % type(obj).__name__)
    out = collections.OrderedDict()
    for k, v in obj.items():
        out[k[len("module."):] if k.startswith("module.") else k] = v
    return out


def load_checkpoint(model, path_or_state, strict=True, map_location="cpu"):
    """Load a reference checkpoint (path to a `.pth`, or an already loaded object) into `model`.
    Returns the (missing, unexpected) key lists; with strict=True any mismatch raises with both lists spelled out."""
    obj = path_or_state
    if not isinstance(obj, dict):
        obj = torch.load(path_or_state, map_location=map_location, weights_only=True)
    sd = _unwrap(obj)
    own = model.state_dict()
    missing = [k for k in own if k not in sd]
    unexpected = [k for k in sd if k not in own]
    # (a 0-d buffer stored as a 1-element vector, e.g. BatchNorm's num_batches_tracked in old files, is the same thing)
    shape = [k for k in sd if k in own and tuple(sd[k].shape) != tuple(own[k].shape) and not (sd[k].numel() == 1 and own[k].numel() == 1)]
    if strict and (missing or unexpected or shape):
        raise RuntimeError("checkpoint does not match the model: missing %s, unexpected %s, shape mismatch %s"
                           % (missing[:8], unexpected[:8], [(k, tuple(sd[k].shape), tuple(own[k].shape)) for k in shape[:8]]))
    model.load_state_dict({k: (v.reshape(own[k].shape) if v.numel() == 1 else v) for k, v in sd.items() if k in own and k not in shape},
                          strict=False)
    return missing, unexpected


def save_checkpoint(model, path):
    """Write `model`'s state_dict the way the reference's `_save_network` does (CPU tensors, plain `torch.save`,
    trainer.py:207-223); a model wrapped in DataParallel / DistributedDataParallel is unwrapped first."""
    m = model.module if hasattr(model, "module") and isinstance(model.module, torch.nn.Module) else model
    sd = collections.OrderedDict((k, v.detach().cpu()) for k, v in m.state_dict().items())
    torch.save(sd, path)
    return path
