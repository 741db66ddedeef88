"""Mint tests/golden_cnn/*.pt by running the REFERENCE's own `GlobalFeatureExtractorCNN`
(/root/reference/src/feature_extractors.py:6-34, executed in place; its `torch_geometric.nn` import is answered by
the harness shim -- the CNN class itself is plain torch) and `reshape_fd_tensor_to_grid`
(/root/reference/src/utils_data.py:125-141) on synthetic fields.

    python -m oracle.ref_harness.make_golden_cnn        # build container only

Each fixture stores the nodal field, the module's state_dict, the mapping tensor (2-D), and what the reference
produced: the grid it fed the CNN, the features [B, C_out], and the parameter gradients of sum(features * w)."""
from __future__ import annotations

import importlib
import os
import sys
import types

import numpy as np
import torch

_REPO = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, _REPO)
_SHIM = os.path.join(os.path.dirname(os.path.abspath(__file__)), "shim")
REFERENCE_SRC = "/root/reference/src"
GOLDEN_DIR = os.path.join(_REPO, "tests", "golden_cnn")


def load_reference_modules():
    if not os.path.isfile(os.path.join(REFERENCE_SRC, "feature_extractors.py")):
        raise RuntimeError("/root/reference is not present: the harness only runs in the build container")
    for p in (REFERENCE_SRC, _SHIM):
        if p in sys.path:
            sys.path.remove(p)
        sys.path.insert(0, p)
    for name in ("wandb", "matplotlib", "matplotlib.pyplot"):          # imported at the top of utils_data.py, unused here
        sys.modules.setdefault(name, types.ModuleType(name))
    sys.modules["matplotlib"].pyplot = sys.modules["matplotlib.pyplot"]
    for n in ("feature_extractors", "utils_data"):
        sys.modules.pop(n, None)
    fe = importlib.import_module("feature_extractors")
    ud = importlib.import_module("utils_data")
    for m in (fe, ud):
        assert os.path.realpath(m.__file__).startswith("/root/reference/"), m.__file__
    return fe, ud


# name, dim, n, B, (mid, out, layers)
CASES = [
    ("cnn2d_6x6_b3", 2, 6, 3, (8, 8, 4)),
    ("cnn2d_30x30_b2", 2, 30, 2, (8, 8, 4)),
    ("cnn2d_9x9_b4_small", 2, 9, 4, (5, 3, 3)),
    ("cnn1d_21_b4", 1, 21, 4, (8, 8, 4)),
    ("cnn1d_200_b3", 1, 200, 3, (8, 8, 4)),
]


def main():
    from g_adaptivity_b200 import synth
    fe, ud = load_reference_modules()
    os.makedirs(GOLDEN_DIR, exist_ok=True)
    for k, (name, dim, n, B, (mid, out, L)) in enumerate(CASES):
        torch.manual_seed(100 + k)
        md = (n, n) if dim == 2 else (n,)
        data = synth.make_batch(md, B, seed=40 + k)
        u = data.f_tensor.clone()                                     # nodal values, mesh-major, [B * N]
        model = fe.GlobalFeatureExtractorCNN(1, mid, out, dim=dim, num_layers=L)
        mapping = None
        if dim == 2:
            ds = synth.SyntheticDataset(2, md)
            mapping = ds.mapping_tensor
            grid = ud.reshape_fd_tensor_to_grid(u, mapping, [n, n], B, 2)
        else:
            grid = ud.reshape_fd_tensor_to_grid(u, None, [n, n], B, 1)
        feats = model(grid.unsqueeze(1))
        w = torch.randn(feats.shape, generator=torch.Generator().manual_seed(7 + k))
        (feats * w).sum().backward()
        fx = {"source": "/root/reference/src/feature_extractors.py:6-34 + utils_data.py:125-141, executed in place",
              "dim": dim, "n": n, "B": B, "channels": [mid, out, L], "u": u, "mapping_tensor": mapping,
              "grid": grid.detach().clone(), "state_dict": {k_: v.detach().clone() for k_, v in model.state_dict().items()},
              "features": feats.detach().clone(), "cotangent": w,
              "grads": {k_: p.grad.detach().clone() for k_, p in model.named_parameters()}}
        torch.save(fx, os.path.join(GOLDEN_DIR, name + ".pt"))
        print(f"{name}: grid {tuple(grid.shape)} features {tuple(feats.shape)} |f|max {feats.abs().max().item():.4f}")


if __name__ == "__main__":
    main()
