"""Mint `tests/golden/*.pt` by running the REFERENCE's own `GNN` / `GRAND_plusConv` /
`GRAND_conv` source (from /root/reference/src, executed in place through
`oracle/ref_harness/load_reference.py`) on synthetic batches.

    python -m oracle.ref_harness.make_golden          # from the repo root, in the build container

Each fixture stores the complete inputs, the reference model's `state_dict`, and what the
reference produced: relocated coordinates `x_phys`, the filtered `edge_index` the conv layers
saw (`stored_ei`), the last layer's attention `stored_alpha`, the L1 mesh loss against
`data.x_phys` and the parameter gradients autograd returned.  The tests compare the oracle
(`oracle/gnn_oracle.py`) and, on the GPU, the CUDA path against these numbers.

What this pins: every line of `src/GNN.py:144-306` and `src/GRAND_plus.py:114-343,366-382`
that runs for the in-scope options.  What it cannot pin: torch-geometric itself, which is
replaced by the shim (restated from PyG 2.4.0's published algorithm).
"""
from __future__ import annotations

import os
import sys

import numpy as np
import torch
import torch.nn.functional as F

_REPO = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
if _REPO not in sys.path:
    sys.path.insert(0, _REPO)

from g_adaptivity_b200 import synth  # noqa: E402
from oracle.ref_harness import load_reference  # noqa: E402

GOLDEN_DIR = os.path.join(_REPO, "tests", "golden")

# name, list of mesh_dims (one per mesh), opt overrides, extras
CASES = [
    ("cfg1_15x15", [[15, 15]], {}, {}),
    ("b3_6x6", [[6, 6]] * 3, {}, {}),
    ("selfloops_7x7", [[7, 7]] * 2, {"self_loops": True}, {}),
    ("temp_fixed_6x6", [[6, 6]] * 2, {"softmax_temp_type": "fixed", "softmax_temp": 2.0}, {}),
    ("learn_step_6x6", [[6, 6]] * 2, {"learn_step": True}, {"steps": [0.05, 0.1, 0.2, 0.15]}),
    ("noshare_6x6", [[6, 6]] * 2, {"share_conv": False, "num_layers": 3}, {}),
    ("nofix_6x6", [[6, 6]] * 2, {"fix_boundary": False}, {}),
    ("normalize_6x6", [[6, 6]] * 2, {"gnn_normalize": True}, {}),
    ("burgers_1d_21x4", [[21]] * 4, "burgers", {}),
    ("poisson_1d_15x2", [[15]] * 2, {}, {}),
    ("mixed_5_7_6", [[5, 5], [7, 7], [6, 6]], {}, {}),
    ("hidden16_6x6", [[6, 6]] * 2, {"hidden_dim": 16}, {}),
    ("hidden2_trunc_6x6", [[6, 6]] * 2, {"hidden_dim": 2}, {}),
    ("scaled_w4_8x8", [[8, 8]] * 2, {}, {"weight_scale": 4.0}),
    ("nofeat_f_6x6", [[6, 6]] * 2, {"gnn_inc_feat_f": False}, {}),
    ("layers8_tau02_6x6", [[6, 6]] * 2, {"num_layers": 8, "time_step": 0.2}, {}),
    # learnable global temperature (GRAND_plus.py:152-154,328-329): the reference leaves the parameter uninitialised;
    # it is assigned here, as loading a checkpoint would
    ("temp_learnable_a_7x7", [[7, 7]] * 3, {"softmax_temp_type": "learnable_a"}, {"sm_temp_a": 1.7}),
    ("temp_learnable_a_noshare_6x6", [[6, 6]] * 2, {"softmax_temp_type": "learnable_a", "share_conv": False,
                                                    "num_layers": 3}, {"sm_temp_a": [0.6, 1.3, 2.5]}),
]


# global CNN features (row f3) through the reference's own forward: 1-D only -- the reference's 2-D branch reshapes to
# [num_nodes, num_nodes] with num_nodes = the NODE count (src/GNN.py:244-245) and raises for every 2-D batch.
# Written to tests/golden_glob/ (own test file: these cases have no streaming / attention read-out variants).
GLOB_DIR = os.path.join(_REPO, "tests", "golden_glob")
GLOB_CASES = [
    ("globfeat_1d_15x3", [[15]] * 3, {"gnn_inc_glob_feat_f": True, "gnn_inc_glob_feat_uu": True}, {}),
    ("globfeat_f_1d_21x2_hidden16", [[21]] * 2, {"gnn_inc_glob_feat_f": True, "hidden_dim": 16}, {}),
    ("globfeat_uu_1d_12x4_noshare", [[12]] * 4, {"gnn_inc_glob_feat_uu": True, "share_conv": False, "num_layers": 3,
                                                  "gnn_inc_feat_f": False}, {}),
]


def build_inputs(mesh_dims_list, burgers: bool, seed: int):
    if len({tuple(m) for m in mesh_dims_list}) == 1:
        return synth.make_batch(mesh_dims_list[0], len(mesh_dims_list), seed=seed, burgers=burgers)
    return synth.make_mixed_batch(mesh_dims_list, seed=seed)


def make_opt(mesh_dims, overrides):
    if overrides == "burgers":
        return synth.burgers_opt(mesh_dims), True
    return synth.default_opt(mesh_dims, **overrides), False


def run_case(gnn_mod, name, mesh_dims_list, overrides, extras, seed=0):
    opt, burgers = make_opt(mesh_dims_list[0], overrides)
    data = build_inputs(mesh_dims_list, burgers, seed)
    dim = len(mesh_dims_list[0])
    dataset = synth.SyntheticDataset(dim, mesh_dims_list[0])
    if dataset.x_comp_shared is None:        # read for its length by the global-feature branch (src/GNN.py:244)
        dataset.x_comp_shared = synth.make_batch(mesh_dims_list[0], 1).x_comp
    torch.manual_seed(opt["seed"])
    with load_reference.quiet():
        model = gnn_mod.GNN(dataset, opt)
    with torch.no_grad():
        if "weight_scale" in extras:
            for n, p in model.named_parameters():
                if "lin_key" in n or "lin_query" in n:
                    p.mul_(extras["weight_scale"])
        if "steps" in extras:
            for p, v in zip(model.steps, extras["steps"]):
                p.fill_(v)
        if "sm_temp_a" in extras:
            vals = extras["sm_temp_a"]
            convs = [model.conv_layers[0]] if opt["share_conv"] else list(model.conv_layers)
            for k, c in enumerate(convs):
                c.sm_temp_a.fill_(vals[k] if isinstance(vals, (list, tuple)) else vals)
    model.train()
    out = model(data)
    target = data.x_phys if data.x_phys.dim() == 2 else data.x_phys.unsqueeze(-1)
    loss = F.l1_loss(out, target)
    loss.backward()
    conv = model.conv_layers[-1]
    grads = {n: p.grad.detach().clone() for n, p in model.named_parameters() if p.grad is not None}
    no_grad = [n for n, p in model.named_parameters() if p.requires_grad and p.grad is None]
    inputs = {k: getattr(data, k).clone() for k in
              ("edge_index", "batch", "x_comp", "x_phys", "f_tensor", "uu_tensor", "u_true_tensor",
               "to_boundary_edge_mask", "to_corner_nodes_mask", "diff_boundary_edges_mask")}
    fixture = {
        "name": name,
        "mesh_dims_list": [list(m) for m in mesh_dims_list],
        "opt_overrides": overrides if isinstance(overrides, dict) else {"__preset__": overrides},
        "extras": extras,
        "seed": seed,
        "inputs": inputs,
        "corner_nodes": [torch.from_numpy(np.asarray(c, dtype=np.int64)) for c in data.corner_nodes],
        "state_dict": {k: v.detach().clone() for k, v in model.state_dict().items()},
        "x_phys": out.detach().clone(),
        "edge_index_filtered": conv.stored_ei.detach().clone(),
        "alpha_last": conv.stored_alpha.detach().clone(),
        "loss": float(loss.item()),
        "grads": grads,
        "params_without_grad": no_grad,
        "reference_files": ["src/GNN.py", "src/GRAND_plus.py", "src/params.py"],
    }
    return fixture


def main():
    gnn_mod, grand_mod, params_mod = load_reference.load()
    os.makedirs(GOLDEN_DIR, exist_ok=True)
    only_glob = "--glob-only" in sys.argv
    only = [a.split("=", 1)[1] for a in sys.argv if a.startswith("--only=")]
    if only:            # mint the named cases without touching the other fixtures
        todo = [(GOLDEN_DIR, c) for c in CASES if c[0] in only] + [(GLOB_DIR, c) for c in GLOB_CASES if c[0] in only]
        for out_dir, (name, mesh_dims_list, overrides, extras) in todo:
            fx = run_case(gnn_mod, name, mesh_dims_list, overrides, extras)
            torch.save(fx, os.path.join(out_dir, name + ".pt"))
            print(f"{name}: loss={fx['loss']:.6e} grads={sorted(fx['grads'])}")
        return
    os.makedirs(GLOB_DIR, exist_ok=True)
    todo = ([] if only_glob else [(GOLDEN_DIR, c) for c in CASES]) + [(GLOB_DIR, c) for c in GLOB_CASES]
    for out_dir, (name, mesh_dims_list, overrides, extras) in todo:
        fx = run_case(gnn_mod, name, mesh_dims_list, overrides, extras)
        path = os.path.join(out_dir, name + ".pt")
        torch.save(fx, path)
        print(f"{name:24s} N={fx['x_phys'].shape[0]:5d} E={fx['edge_index_filtered'].shape[1]:6d} "
              f"loss={fx['loss']:.6e} grads={sorted(fx['grads'])[:2]}.. -> {os.path.relpath(path, _REPO)}")


if __name__ == "__main__":
    main()
