"""Shared helpers for the parity tests: golden-fixture loading and model construction."""
import glob
import os

import numpy as np
import torch

from g_adaptivity_b200 import synth

GOLDEN_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def golden_names():
    return sorted(os.path.splitext(os.path.basename(p))[0] for p in glob.glob(os.path.join(GOLDEN_DIR, "*.pt")))


def load_golden(name):
    return torch.load(os.path.join(GOLDEN_DIR, name + ".pt"), weights_only=True)


def fixture_opt(fx):
    ov = fx["opt_overrides"]
    md = fx["mesh_dims_list"][0]
    if ov.get("__preset__") == "burgers":
        return synth.burgers_opt(md)
    return synth.default_opt(md, **ov)


def fixture_batch(fx):
    """Rebuild the `Batch` the reference consumed from the tensors stored in the fixture."""
    b = synth.Batch()
    for k, v in fx["inputs"].items():
        setattr(b, k, v.clone())
    b.corner_nodes = [c.numpy().copy() for c in fx["corner_nodes"]]
    b.pde_params = {"centers": [], "scales": []}
    b._num_graphs = len(fx["mesh_dims_list"])
    b.mesh_sizes = [int(np.prod(m)) for m in fx["mesh_dims_list"]]
    return b


def fixture_dataset(fx):
    md = fx["mesh_dims_list"][0]
    return synth.SyntheticDataset(len(md), md)


def rel_err(a, b):
    """max |a-b| / max |b| (the parity metric of SURVEY section 4, item 3)."""
    a, b = a.detach().double().cpu(), b.detach().double().cpu()
    denom = b.abs().max().item()
    return (a - b).abs().max().item() / (denom if denom > 0 else 1.0)
