"""Inference forward of the cls network (BASELINE configs[1], 32 clouds): wall time, kernel-class times and the time
of every conv module (CUDA events around the module calls).  A/B knobs travel through the environment
(EPN_FU_HALVES=0, EPN_FUSED=0, ...).   python tools/fwd_prof.py [batch]"""
import ctypes
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from bench import CLASSES, synthetic_clouds  # noqa: E402
from epn_pointcloud_b200 import _lib  # noqa: E402
from epn_pointcloud_b200.heads import ClsSO3ConvModel, cls_model_params  # noqa: E402
from epn_pointcloud_b200.modules import InterSO3Conv, IntraSO3Conv  # noqa: E402

B = int(sys.argv[1]) if len(sys.argv) > 1 else 32
L = _lib.lib()
m = ClsSO3ConvModel(cls_model_params(1024, 60)).cuda().train()
x = synthetic_clouds(B, 1024, 2).cuda()
events = {}


def pre(name):
    def f(mod, inp):
        e = torch.cuda.Event(enable_timing=True)
        e.record()
        events.setdefault(name, []).append([e, None])
    return f


def post(name):
    def f(mod, inp, out):
        e = torch.cuda.Event(enable_timing=True)
        e.record()
        events[name][-1][1] = e
    return f


with torch.no_grad():
    for _ in range(3):
        m(x)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10):
        m(x)
    e1.record()
    torch.cuda.synchronize()
    print("fwd ms", e0.elapsed_time(e1) / 10)
    L.epn_profile_enable(1)
    m(x)
    torch.cuda.synchronize()
    L.epn_profile_enable(0)
    ms = (ctypes.c_double * len(CLASSES))()
    n = (ctypes.c_longlong * len(CLASSES))()
    L.epn_profile_read(ctypes.cast(ms, ctypes.c_void_p), ctypes.cast(n, ctypes.c_void_p), len(CLASSES))
    print({c: (round(ms[i], 2), int(n[i])) for i, c in enumerate(CLASSES)}, "sum", round(sum(ms), 2))
    hs = []
    for name, mod in m.named_modules():
        if isinstance(mod, (InterSO3Conv, IntraSO3Conv)):
            hs.append(mod.register_forward_pre_hook(pre(name)))
            hs.append(mod.register_forward_hook(post(name)))
    for _ in range(3):
        m(x)
    torch.cuda.synchronize()
    for name, ev in events.items():
        print("%-40s %.3f ms" % (name, sum(a.elapsed_time(b) for a, b in ev) / len(ev)))
