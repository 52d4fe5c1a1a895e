import json
import os
import sys

import numpy as np
import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


def pytest_collection_modifyitems(config, items):
    if torch.cuda.is_available():
        return
    skip = pytest.mark.skip(reason="no CUDA device")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)


class Golden(dict):
    """npz fixture -> dict of torch tensors; 'sd.*' / 'grad.*' entries are grouped."""

    def state_dict(self, prefix="sd."):
        return {k[len(prefix):]: v for k, v in self.items() if k.startswith(prefix)}

    def grads(self, prefix="grad."):
        return {k[len(prefix):]: v for k, v in self.items() if k.startswith(prefix)}


def load_golden(name):
    out = Golden()
    with np.load(os.path.join(GOLDEN, name + ".npz")) as d:
        for k in d.files:
            a = d[k]
            if a.dtype.kind in "US":
                out[k] = json.loads(str(a))
            else:
                out[k] = torch.from_numpy(np.ascontiguousarray(a))
    return out


def rel_err(a, b):
    """max |a-b| / max |b|  -- the 'relative fp32' measure used for every feature tensor."""
    a, b = a.detach().double().cpu(), b.detach().double().cpu()
    return float((a - b).abs().max() / b.abs().max().clamp_min(1e-30))


@pytest.fixture(scope="session")
def oracle_lib():
    from oracle import epn_oracle as O
    O.lib()
    return O
