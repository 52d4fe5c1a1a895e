"""epn_pointcloud_b200 -- B200 (sm_100a) engine for the SE(3) separable point convolution
hot path of nintendops/EPN_PointCloud, behind the reference's own op / module surface.

    ops         tensor bindings of the C ABI (libepn_b200.so) + vgtk.cuda.* namespaces
    functional  mirror of vgtk.{pc.sample, spconv.functional, so3conv.functional} for the path
    modules     BasicSO3Conv / InterSO3Conv / IntraSO3Conv / SphericalPointCloud
    blocks      Inter/Intra/Separable/Basic SO3ConvBlock, preprocess_input, backbone builder
    heads       output heads + full models of the three shipped tasks (classification, 3DMatch, rotation)
    losses      their training losses and the rotation utilities (host-side torch)
    parallel    batch sharding + flat-gradient all-reduce (one process per GPU), CUDA-graph training step

The compute path has no CPU / PyTorch fallback: importing is cheap, but the first op call
raises if libepn_b200.so is missing (build it with `__graft_entry__.build()`).
"""
from . import _lib  # noqa: F401
from . import ops, functional, modules, blocks, heads, losses  # noqa: F401
from .modules import BasicSO3Conv, InterSO3Conv, IntraSO3Conv, SphericalPointCloud  # noqa: F401

__version__ = "0.1.0"
