"""g_adaptivity_b200 -- the GNN mesh-deformer hot path of g-adaptivity on B200 (sm_100a).

Only what the path needs: `csrc/` (CUDA kernels + the C ABI of `include/gadapt.h`), the ctypes
binding, the cached mesh graph, and the host-side mirror of the reference interface
(`GNN`, `GRAND_plusConv`, `GRAND_conv`, `params`).  Importing the package does not need a GPU;
running any kernel does, and raises otherwise -- there is no CPU fallback.
"""
from . import params  # noqa: F401
from .GNN import GNN, build_conv_list, get_conv, get_dec, get_enc, get_nonlin  # noqa: F401
from .GRAND_plus import GRAND_conv, GRAND_plusConv  # noqa: F401
from .graph import GraphCache, MeshGraph, plan_tiles  # noqa: F401

__all__ = ["GNN", "GRAND_plusConv", "GRAND_conv", "MeshGraph", "GraphCache", "plan_tiles", "params",
           "get_conv", "get_enc", "get_dec", "get_nonlin", "build_conv_list"]
