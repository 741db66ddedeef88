"""Config surface of the deformer: the same `opt` keys, CLI flags and defaults as
`src/params.py:199-303` (`get_params`), the string->bool sweep of `:164-177` and the experiment
presets of `:8-161` (`run_params`), so `run_pipeline.py`-style drivers can build `opt` unchanged.

Only the keys the deformer reads change behaviour here ("config / flags"); the
rest are carried so downstream reference code finds them.  New optional keys (default =
reference behaviour):

    ode_method        'euler' (reference, src/GNN.py:288-291) | 'rk4' (extension: fused RK4 steps, forward and hand-written backward through GNN.forward)
    gad_tile_nodes    nodes per CTA tile for the mesh-resident kernels (default 1024)
    gad_force_stream  force the per-layer streaming kernels
    gad_store_alpha   keep what `conv.stored_alpha` needs after each forward (default True)
    gad_sync_timestamp  synchronise the stream before stamping `model.end_MLmodel`
"""
from __future__ import annotations

import argparse
import ast
import os
import random

import numpy as np
import torch

# (flag, type, default, choices)  -- one row per `parser.add_argument` of the reference
_CLI = [
    # data
    ("dataset", str, "grid", ["fd_mmpde_1d", "fd_mmpde_2d", "fd_ma_2d", "grid", "noisey_grid", "triangles"]),
    ("data_type", str, "randg", ["all", "structured", "randg", "randg_mix"]),
    ("fast_M2N_monitor", str, "slow", ["fast", "slow", "superslow"]),
    ("M2N_alpha", float, None, None), ("M2N_beta", float, None, None),
    ("mesh_type", str, "ma", ["mmpde", "ma", "M2N"]),
    ("data_name", str, "test", None), ("data_train_test", str, "train", ["train", "test"]),
    ("num_train", int, 100, None), ("num_test", int, 25, None),
    ("train_frac", float, None, None), ("test_frac", float, None, None),
    # mesh
    ("fix_boundary", str, "True", None), ("mon_reg", float, 0.1, None), ("mon_power", float, 0.2, None),
    # pde
    ("pde_type", str, "Poisson", ["Poisson", "Burgers"]), ("boundary", str, "dirichlet", None),
    ("num_gauss", int, 1, None), ("rand_gauss", bool, False, None), ("scale", float, 0.2, None),
    ("center", float, 0.5, None),
    # fem
    ("eval_quad_points", int, 101, None), ("stiff_quad_points", int, 3, None), ("load_quad_points", int, 101, None),
    # model
    ("model", str, "GNN", ["fixed_mesh_1D", "fixed_mesh_2D", "backFEM_1D", "backFEM_2D", "GNN", "MLP"]),
    ("num_layers", int, 4, None), ("hidden_dim", int, 8, None), ("global_feat_dim", int, 8, None),
    ("enc", str, "identity", ["identity", "lin_layer", "mlp"]), ("dec", str, "identity", ["identity", "lin_layer", "mlp"]),
    ("non_lin", str, "identity", ["identity", "relu", "tanh", "sigmoid", "leaky_relu"]),
    ("residual", str, "True", None), ("mesh_params", str, "internal", None), ("time_step", float, 0.1, None),
    # GNN
    ("conv_type", str, "GCN", ["GCN", "GAT", "GRAND", "GRAND_plus", "GAT_plus", "Laplacian"]),
    ("share_conv", str, "True", None), ("gnn_inc_feat_f", str, "True", None), ("gnn_inc_feat_uu", str, "False", None),
    ("gnn_inc_glob_feat_f", str, "True", None), ("gnn_inc_glob_feat_uu", str, "True", None),
    ("gnn_normalize", str, "False", None),
    # regularisation
    ("self_loops", str, "False", None), ("softmax_temp_type", str, None, ["none", "fixed", "learnable"]),
    ("softmax_temp", float, 2.0, None), ("learn_step", str, "False", None), ("gnn_dont_train", str, "False", None),
    ("reg_skew", str, "False", None),
    ("gat_plus_type", str, "GAT_res_lap", ["GAT_res_lap", "GAT_lin", "GAT", "GAT_phys", "None"]),
    # Burgers
    ("gauss_amplitude", float, 0.25, None), ("burgers_limits", float, 3.0, None),
    ("plots_multistep_eval", str, "False", None), ("plots_mesh_movement", str, "False", None),
    # training
    ("seed", int, 42, None), ("device", str, "cpu", None), ("batch_size", int, 1, None), ("epochs", int, 100, None),
    ("lr", float, 0.001, None), ("dropout", float, 0.0, None), ("decay", float, 0.0, None),
    ("loss_type", str, "mesh_loss", ["mesh_loss", "pde_loss", "pinn_loss"]), ("loss_fn", str, "l1", ["mse", "l1"]),
    ("solver", str, "torch_FEM", ["firedrake", "torch_FEM", "BVP"]),
    ("evaler", str, "analytical", ["fd_fine", "fd_coarse", "analytical"]),
    # plots
    ("show_plots", str, "True", None), ("show_dataset_plots", str, "True", None),
    ("show_train_evol_plots", str, "True", None), ("show_mesh_evol_plots", str, "True", None),
    ("show_mesh_plots", str, "False", None),
    # B200 extensions
    ("ode_method", str, "euler", ["euler", "rk4"]),
]
# list-valued flags (`nargs='+'`)
_CLI_LISTS = [
    ("mesh_dims_train", [[15, 15], [20, 20]]), ("mesh_dims_test", [[i, i] for i in range(12, 24, 1)]),
    ("num_gauss_range", [1, 2, 3, 5, 6]), ("mesh_dims", [10, 10]), ("overfit_num", None),
]


def get_params(argv=None) -> dict:
    parser = argparse.ArgumentParser()
    for name, typ, default, choices in _CLI:
        kw = {"type": typ, "default": default}
        if choices is not None:
            kw["choices"] = choices
        parser.add_argument("--" + name, **kw)
    for name, default in _CLI_LISTS:
        parser.add_argument("--" + name, nargs="+", default=default)
    return vars(parser.parse_args(argv))


def t_or_f(v):
    if isinstance(v, bool):
        return v
    if v in ("True", "true"):
        return True
    if v in ("False", "false"):
        return False
    return v


def tf_sweep_args(opt: dict) -> dict:
    for k in list(opt):
        opt[k] = t_or_f(opt[k])
    return opt


def get_arg_list(arg_list):
    """`mesh_dims` arrives as a list of ints or as one string holding a python list
    (`src/params.py:190-196`)."""
    if type(arg_list[0]) == int:
        return arg_list
    # the reference calls eval() on the CLI string; a literal parser accepts the same inputs ("[15, 15]") and
    # nothing else
    return list(ast.literal_eval(arg_list[0]))


def set_seed(seed: int = 42) -> None:
    np.random.seed(seed)
    random.seed(seed)
    torch.manual_seed(seed)
    if torch.cuda.is_available():
        torch.cuda.manual_seed(seed)
    torch.backends.cudnn.deterministic = True
    torch.backends.cudnn.benchmark = False
    os.environ["PYTHONHASHSEED"] = str(seed)


def run_params(opt: dict) -> dict:
    """The reference's hard-coded experiment preset (`src/params.py:8-161`): values are written
    over whatever the CLI produced.  Poisson 2-D, GNN model, GRAND_plus deformer by default;
    `opt['pde_type'] = 'Burgers'` before the call selects the 1-D Burgers preset."""
    burgers = opt.get("pde_type") == "Burgers"
    opt.setdefault("pde_type", "Poisson")
    opt["data_type"] = "randg"
    if burgers:
        opt.update(mesh_type="mmpde", dataset="fd_mmpde_1d", mesh_dims=[15], mon_reg=0.1, num_gauss=1)
    else:
        opt.update(mesh_type="ma", dataset="fd_ma_2d", mesh_dims=[11, 11], mon_reg=0.01)
    opt["model"] = "GNN"
    opt.update(num_gauss=2, rand_gauss=True, num_train=25, num_test=25)
    opt.update(fix_boundary=True, eval_quad_points=101, stiff_quad_points=3, load_quad_points=101)
    opt.update(epochs=1, gnn_dont_train=False, loss_type="pde_loss", loss_fn="l1", solver="torch_FEM",
               gnn_inc_feat_f=True, gnn_inc_feat_uu=True, gnn_inc_glob_feat_f=False, gnn_inc_glob_feat_uu=False,
               gnn_normalize=False, conv_type="GRAND_plus", gat_plus_type="GAT_res_lap", enc="identity",
               dec="identity", residual=True, share_conv=True, non_lin="identity", num_layers=4, time_step=0.1,
               hidden_dim=8, global_feat_dim=8, lr=0.001)
    if burgers:
        opt.update(gauss_amplitude=0.25, burgers_limits=3.0, num_train=20, num_test=5, scale=0.1, mon_reg=0.1,
                   num_gauss=1, loss_type="modular", mesh_dims=[21], conv_type="GRAND",
                   grad_type="burgers_timestep_loss_direct_mse", epochs=100, global_feat_dim=8,
                   num_fine_mesh_points=40, gnn_inc_feat_f=False, tau=1 / 20.0, nu=0.001, num_time_steps=1,
                   num_eval_time_steps=20)
    return opt
