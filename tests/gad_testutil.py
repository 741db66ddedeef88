"""Shared helpers for the parity tests: golden-fixture loading and model construction."""
import glob
import os

import copy

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


def fp64_grads_and_noise_floor(ds, opt, data, ref, ref_grads):
    """Gradients of the oracle evaluated in fp64, and per parameter the relative deviation of the
    fp32 oracle from them.  On ill-conditioned cases (1-D Burgers batches: gradients ~1e-7 after
    cancellation over 10^4 nodes) the fp32 oracle itself is 1e-3 away from the exact gradient, so
    the parity bar there is max(1e-4, 4 x that noise floor) against the fp64 values: the kernels use
    the MUFU approximations ex2 / lg2 / rcp (2^-22 relative error, i.e. 4 ulp) where the CPU oracle
    has <= 1-ulp libm calls, so up to 4 x the fp32 oracle's own rounding noise is expected."""
    from oracle import gnn_oracle
    ref64 = gnn_oracle.GNNRef(ds, copy.deepcopy(opt))
    ref64.load_state_dict(ref.state_dict())
    ref64 = ref64.double()
    d64 = data.clone()
    for k in d64.keys():
        v = getattr(d64, k)
        if torch.is_tensor(v) and v.dtype == torch.float32:
            setattr(d64, k, v.double())
    gnn_oracle.mesh_loss(ref64(d64), d64.x_phys).backward()
    g64 = {n: p.grad for n, p in ref64.named_parameters() if p.grad is not None}
    scale = max(g.abs().max().item() for n, g in g64.items() if "lin_key.bias" not in n)
    floor = {}
    for n, g in g64.items():
        denom = max(g.abs().max().item(), 1e-3 * scale, 1e-300)
        floor[n] = (ref_grads[n].double() - g).abs().max().item() / denom
    return g64, floor, scale


def check_grads_conditioned(got, g64, floor, scale, tol=1e-4):
    for n, g in g64.items():
        if "lin_key.bias" in n:
            assert got[n].abs().max().item() <= 1e-4 * scale, n
            continue
        denom = max(g.abs().max().item(), 1e-3 * scale)
        err = (got[n].double() - g).abs().max().item() / denom
        assert err <= max(tol, 4.0 * floor[n]), (n, err, floor[n])
