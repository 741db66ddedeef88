"""TEST INFRASTRUCTURE -- CPU oracle for the g-adaptivity deformer hot path.

A pure-PyTorch, CPU, op-for-op restatement of

    src/GNN.py:72-90     identity encoder
    src/GNN.py:144-188   GNN.__init__
    src/GNN.py:190-306   GNN.forward (mesh_loss / modular return)
    src/GRAND_plus.py:114-188, 204-343   GRAND_plusConv
    src/GRAND_plus.py:366-382            GRAND_conv (PyG TransformerConv, identity value)

on top of the torch_geometric 2.4.0 primitives restated in `oracle/pyg_semantics.py`.

Only `tests/`, `__graft_entry__.smoke()` and `bench.py`'s cpu_baseline / `--impl reference`
legs may import this module, and only as the checker or the timed CPU baseline.  The product
package `g_adaptivity_b200` never imports it and has no CPU fallback.

Parity status.  The reference ships no tests or golden vectors and cannot be
imported whole (PyG, Firedrake, torchquad are absent).  The oracle is pinned three ways:
  1. against the reference's OWN `GRAND_plus.py` / `GNN.py` source, executed in this container
     through the PyG shim of `oracle/ref_harness/` -> committed fixtures `tests/golden/*.pt`
     (generator: `oracle/ref_harness/make_golden.py`);
  2. against the dense-matrix formulation `dense_layer` below (fp64, 1e-12);
  3. by the structural invariants of SURVEY section 4 (row-stochastic alpha, fixed corners, ...).
What stays unpinned is PyG itself (restated from its published 2.4.0 algorithm).

Extension (not in the reference): `opt['ode_method'] == 'rk4'` turns every layer's explicit
Euler update into one classical RK4 step of the same vector field F(y) = A(y) y - y with step
`time_step` (form of `classical_meshing/ma_mesh_1d.py:65-70`); BASELINE config 4 uses it.
"""
from __future__ import annotations

import math
import time
from typing import List, Optional, Tuple

import numpy as np
import torch
import torch.nn as nn
import torch.nn.functional as F
from torch import Tensor

from . import pyg_semantics as pyg


# --------------------------------------------------------------------------------------
# layers
# --------------------------------------------------------------------------------------
class LinearRef(nn.Module):
    """torch_geometric.nn.dense.linear.Linear: y = x W^T + b, W [out, in]; default init is
    kaiming-uniform(a=sqrt(5)) for W and U(-1/sqrt(in), 1/sqrt(in)) for b, i.e. nn.Linear's."""

    def __init__(self, in_channels: int, out_channels: int, bias: bool = True):
        super().__init__()
        self.weight = nn.Parameter(torch.empty(out_channels, in_channels))
        self.bias = nn.Parameter(torch.empty(out_channels)) if bias else None
        self.reset_parameters()

    def reset_parameters(self):
        nn.init.kaiming_uniform_(self.weight, a=math.sqrt(5))
        if self.bias is not None:
            bound = 1.0 / math.sqrt(self.weight.shape[1])
            nn.init.uniform_(self.bias, -bound, bound)

    def forward(self, x: Tensor) -> Tensor:
        return F.linear(x, self.weight, self.bias)


class GRANDPlusConvRef(nn.Module):
    """`GRAND_plusConv` as `get_conv` builds it (`src/GNN.py:117-119`): heads=1, concat=False,
    beta=False, dropout=0, edge_dim=None, bias=False, root_weight=False."""

    def __init__(self, opt: dict, in_channels: int, out_channels: int, heads: int = 1, **kwargs):
        super().__init__()
        assert heads == 1
        self.opt = opt
        self.in_channels, self.out_channels, self.heads = in_channels, out_channels, heads
        self.lin_key = LinearRef(in_channels, heads * out_channels)      # GRAND_plus.py:146
        self.lin_query = LinearRef(in_channels, heads * out_channels)    # GRAND_plus.py:147
        self.lin_skip = LinearRef(in_channels, out_channels, bias=False)  # :178 (never used, root_weight=False)
        if opt.get("softmax_temp_type") == "learnable_v":
            raise NotImplementedError("softmax_temp_type='learnable_v' applies Linear(C, H) to the [E, H] logits "
                                      "(GRAND_plus.py:158,331): a shape error in the reference")
        if opt.get("softmax_temp_type") == "learnable_a":
            # :152-154 -- `nn.Parameter(torch.Tensor(1, heads, 1))`, i.e. UNINITIALISED memory in the reference; usable
            # once a value is loaded / assigned.  Initialised to opt['softmax_temp'] here.
            self.sm_temp_a = nn.Parameter(torch.full((1, heads, 1), float(opt.get("softmax_temp", 1.0))))
        if opt.get("reg_skew"):
            raise NotImplementedError("reg_skew needs a Firedrake mesh (GRAND_plus.py:280-324)")
        self.stored_ei = None
        self.stored_alpha = None
        self._always_store = False

    def forward(self, x: Tensor, edge_index: Tensor, global_features=None, mesh=None) -> Tensor:
        H, C = self.heads, self.out_channels
        query = self.lin_query(x).view(-1, H, C)                         # :225
        key = self.lin_key(x).view(-1, H, C)                             # :226
        value = x.view(-1, H, C)                                         # :227 (identity)
        store = {}

        def message(query_i, key_j, value_j, index, size_i, **_):
            alpha = (query_i * key_j).sum(dim=-1) / math.sqrt(self.out_channels)   # :279
            if self.opt.get("softmax_temp_type") == "fixed":                        # :326-327
                alpha = pyg.softmax(alpha / self.opt["softmax_temp"], index, None, size_i)
            elif self.opt.get("softmax_temp_type") == "learnable_a":                # :328-329
                alpha = pyg.softmax(alpha / self.sm_temp_a.squeeze(2), index, None, size_i)
            else:                                                                   # :332-333
                alpha = pyg.softmax(alpha, index, None, size_i)
            store["alpha"] = alpha
            alpha = F.dropout(alpha, p=0.0, training=self.training)                 # :336
            return value_j * alpha.view(-1, self.heads, 1)                          # :338-343

        out = pyg.propagate_add(edge_index, message, x.shape[0], query=query, key=key, value=value)  # :233
        out = out.mean(dim=1)                                            # :242 (concat=False)
        if self._always_store or isinstance(self.opt.get("show_mesh_evol_plots"), bool):   # :253-256
            self.stored_ei = edge_index
            self.stored_alpha = store["alpha"]
        return out - x                                                   # :267


class GRANDConvRef(GRANDPlusConvRef):
    """`GRAND_conv` (`src/GRAND_plus.py:366-382`): PyG TransformerConv(heads=1, concat=False,
    root_weight=False, bias=False) with identity value -- the same arithmetic, no temperature.
    (TransformerConv has no `lin_skip` bias either; its `lin_skip` Linear exists with bias=False.)"""

    def __init__(self, opt: dict, in_channels: int, out_channels: int, heads: int = 1, **kwargs):
        opt_plain = dict(opt)
        opt_plain["softmax_temp_type"] = None
        opt_plain["reg_skew"] = False
        super().__init__(opt_plain, in_channels, out_channels, heads)
        self.opt = opt_plain
        self._always_store = True   # return_attention_weights=True (:381) stores ei/alpha every call

    def forward(self, x: Tensor, edge_index: Tensor) -> Tensor:
        return super().forward(x, edge_index)


# --------------------------------------------------------------------------------------
# graph prologue (src/GNN.py:206-223) and feature assembly (:225-239)
# --------------------------------------------------------------------------------------
def get_arg_list(arg_list):
    """`src/params.py:190-196` without the print."""
    if type(arg_list[0]) == int:
        return arg_list
    return eval(arg_list[0])


def filtered_edge_index(data, opt: dict, dim: int) -> Tensor:
    """The `edge_index` the conv layers see (`src/GNN.py:194,206-223`)."""
    batch = data.batch
    num_in_batch = int(batch.max().item()) + 1
    edge_index = data.edge_index
    if opt["fix_boundary"]:
        mask = ~data.to_boundary_edge_mask * ~data.to_corner_nodes_mask * ~data.diff_boundary_edges_mask
        edge_index = edge_index[:, mask]
        if dim == 1:
            n = opt["mesh_dims"][0]
            corner_nodes = torch.cat([torch.tensor([0 + b * n, (1 + b) * n - 1])
                                      for b in range(num_in_batch)]).repeat(2, 1)
            edge_index = torch.cat([edge_index, corner_nodes], dim=1)
        elif dim == 2:
            corner_nodes = torch.stack([torch.from_numpy(np.asarray(arr)) for arr in data.corner_nodes])
            num_each_nodes = batch.unique(return_counts=True)[1]
            cum_num_each_nodes = torch.cumsum(num_each_nodes, dim=0)
            corner_nodes[1:] += cum_num_each_nodes[:-1].unsqueeze(-1)
            corner_edges = corner_nodes.reshape(-1).repeat(2, 1)
            edge_index = torch.cat([edge_index, corner_edges], dim=1)
    if opt["self_loops"]:
        num_nodes = data.x_comp.size(0)
        edge_index, _ = pyg.remove_self_loops(edge_index)
        edge_index, _ = pyg.add_self_loops(edge_index, num_nodes=num_nodes)
    return edge_index


def assemble_features(data, opt: dict, dim: int) -> Tensor:
    """`src/GNN.py:225-239`: [x_comp | f | uu], optional division by the batch-wide signed max."""
    x_comp = data.x_comp
    if dim == 1:
        x_comp = x_comp.unsqueeze(-1)
    features = x_comp
    if opt["gnn_inc_feat_f"]:
        f = data.f_tensor
        if opt["gnn_normalize"]:
            f = f / torch.max(f)
        features = torch.cat([features, f.unsqueeze(-1)], dim=1)
    if opt["gnn_inc_feat_uu"]:
        uu = data.uu_tensor
        if opt["gnn_normalize"]:
            uu = uu / torch.max(uu)
        features = torch.cat([features, uu.unsqueeze(-1)], dim=1)
    return features


def identity_encoder(in_dim: int, out_dim: int) -> nn.Linear:
    """`get_enc(opt['enc']=='identity')` (`src/GNN.py:75-90`): frozen 0/1 pad / truncate."""
    lin = nn.Linear(in_dim, out_dim, bias=False)
    w = torch.zeros(out_dim, in_dim)
    k = min(in_dim, out_dim)
    w[:k, :k] = torch.eye(k)
    lin.weight.data = w
    lin.weight.requires_grad = False
    return lin


class GlobalCNNRef(nn.Module):
    """`GlobalFeatureExtractorCNN` (`src/feature_extractors.py:6-34`): same sub-modules, forward by the restatement
    in oracle/cnn_oracle.py."""

    def __init__(self, in_channels, mid_channels, out_channels, dim=2, num_layers=4):
        super().__init__()
        conv = nn.Conv1d if dim == 1 else nn.Conv2d
        self.convs = nn.ModuleList([conv(in_channels, mid_channels, kernel_size=3, stride=1, padding=1)])
        for _ in range(num_layers - 2):
            self.convs.append(conv(mid_channels, mid_channels, kernel_size=3, stride=1, padding=1))
        self.convs.append(conv(mid_channels, out_channels, kernel_size=3, stride=1, padding=1))

    def forward(self, u):
        from oracle import cnn_oracle
        return cnn_oracle.cnn_features(u, [c.weight for c in self.convs], [c.bias for c in self.convs])


class GNNRef(nn.Module):
    """`GNN` (`src/GNN.py:144-306`) for the in-scope option set."""

    def __init__(self, dataset, opt: dict):
        super().__init__()
        self.dataset, self.opt = dataset, opt
        self.dim = dataset.num_x_comp_features                       # :149
        self.mesh_dims = get_arg_list(opt["mesh_dims"])
        self.in_dims = [self.dim]
        if opt["gnn_inc_feat_f"]:
            self.in_dims += [1]
        if opt["gnn_inc_feat_uu"]:
            self.in_dims += [1]
        if opt["gnn_inc_glob_feat_f"]:
            self.in_dims += [opt["global_feat_dim"]]                 # :156-157
        if opt["gnn_inc_glob_feat_uu"]:
            self.in_dims += [opt["global_feat_dim"]]                 # :158-159
        opt["hidden_dims_list"] = self.in_dims                       # :161
        in_dim, hid = sum(self.in_dims), opt["hidden_dim"]
        if opt["enc"] != "identity":
            raise NotImplementedError("enc != identity crashes in the reference (GNN.py:101-105,298)")
        self.enc = identity_encoder(in_dim, hid)                     # :167
        conv_cls = {"GRAND_plus": GRANDPlusConvRef, "GRAND": GRANDConvRef}[opt["conv_type"]]
        layers = []
        shared = conv_cls(opt, hid, hid, heads=1) if opt["share_conv"] else None   # :131-132
        for _ in range(opt["num_layers"]):
            layers.append(shared if shared is not None else conv_cls(opt, hid, hid, heads=1))
        self.conv_layers = nn.ModuleList(layers)                     # :168
        if opt["non_lin"] != "identity" and opt["conv_type"] == "GRAND":
            self.non_lin = {"relu": nn.ReLU(), "tanh": nn.Tanh(), "sigmoid": nn.Sigmoid(),
                            "leaky_relu": nn.LeakyReLU(), "elu": nn.ELU(), "selu": nn.SELU()}[opt["non_lin"]]
        else:
            self.non_lin = nn.Identity()
        self.dec = nn.Identity()                                     # :170
        if opt["gnn_inc_glob_feat_f"]:                               # :172-174
            self.global_feature_extractor_cnn_f = GlobalCNNRef(1, hid, opt["global_feat_dim"], dim=self.dim)
        if opt["gnn_inc_glob_feat_uu"]:                              # :175-177
            self.global_feature_extractor_cnn_uu = GlobalCNNRef(1, hid, opt["global_feat_dim"], dim=self.dim)
        if opt["learn_step"]:                                        # :179-180
            self.steps = nn.ParameterList([nn.Parameter(torch.tensor([opt["time_step"]]))
                                           for _ in range(opt["num_layers"])])
        self.end_MLmodel = None

    def _append_global_features(self, data, features):
        """`src/GNN.py:242-268`: per mesh, the CNN features of the f / uu grid, repeated onto the mesh's nodes and
        appended.  REFERENCE QUIRK: for `data_type != 'randg_mix'` the reference reshapes to
        `[num_nodes, num_nodes]` with `num_nodes = dataset.x_comp_shared.shape[0]` (the NODE count, :244-245), which
        cannot be reshaped to and raises for every 2-D batch; its `randg_mix` branch uses
        `int(sqrt(data.x_comp.shape[0]))` (:247), the side length only for a batch of one.  The intended grid -- one
        n x n image per mesh, what that branch computes for a single mesh -- is what is restated here; the 1-D
        branch (`reshape(batch_size, -1)`, utils_data.py:126-129) is followed as written."""
        opt = self.opt
        if not (opt["gnn_inc_glob_feat_f"] or opt["gnn_inc_glob_feat_uu"]):
            return features
        from oracle import cnn_oracle
        batch = data.batch
        B = int(batch.max().item()) + 1
        repeats = torch.bincount(batch)
        n = int(round((data.x_comp.shape[0] // B) ** 0.5)) if self.dim == 2 else data.x_comp.shape[0] // B
        mapping = getattr(self.dataset, "mapping_tensor", None)
        for name, on, field, inc in (("f", opt["gnn_inc_glob_feat_f"], data.f_tensor, opt["gnn_inc_feat_f"]),
                                     ("uu", opt["gnn_inc_glob_feat_uu"], data.uu_tensor, opt["gnn_inc_feat_uu"])):
            if not on:
                continue
            if inc and opt["gnn_normalize"]:
                field = field / torch.max(field)                     # f / uu were normalised in place above (:231-237)
            grid = cnn_oracle.reshape_fd_tensor_to_grid(field, mapping, [n, n], B, self.dim)
            g = getattr(self, f"global_feature_extractor_cnn_{name}")(grid.unsqueeze(1))
            # (.float() in the reference, :253,266; the input dtype here so that the fp64 cross-checks can run)
            features = torch.cat([features, g.repeat_interleave(repeats, dim=0)], dim=-1).to(features.dtype)
        return features

    def _vector_field(self, layer, x, edge_index, features):
        if self.opt["conv_type"] == "GRAND_plus":
            return layer(x, edge_index, features, None)              # :282
        res = layer(x, edge_index)                                   # :284
        res = F.dropout(res, self.opt["dropout"], training=self.training)
        return self.non_lin(res)                                     # :285-286

    def forward(self, data, return_states: bool = False):
        opt = self.opt
        if opt.get("dropout", 0.0) != 0.0:
            raise NotImplementedError("dropout > 0 is off the hot path")
        if not opt["residual"]:
            raise NotImplementedError("residual=False is only meaningful for non-GRAND convs")
        edge_index = filtered_edge_index(data, opt, self.dim)        # :194,206-223
        features = assemble_features(data, opt, self.dim)            # :225-239
        features = self._append_global_features(data, features)      # :242-268
        x = self.enc(features)                                       # :270
        x = F.dropout(x, opt["dropout"], training=self.training)     # :271
        states = [x]
        method = opt.get("ode_method", "euler")
        for i, layer in enumerate(self.conv_layers):                 # :273
            tau = self.steps[i] if opt["learn_step"] else opt["time_step"]   # :288-291
            if method == "euler":
                res = self._vector_field(layer, x, edge_index, features)
                x = x + tau * res
            elif method == "rk4":
                k1 = self._vector_field(layer, x, edge_index, features)
                k2 = self._vector_field(layer, x + (0.5 * tau) * k1, edge_index, features)
                k3 = self._vector_field(layer, x + (0.5 * tau) * k2, edge_index, features)
                k4 = self._vector_field(layer, x + tau * k3, edge_index, features)
                x = x + (tau / 6.0) * (k1 + 2.0 * k2 + 2.0 * k3 + k4)
            else:
                raise ValueError(method)
            states.append(x)
        x = self.dec(x)                                              # :298
        x_phys = x[:, :self.dim]                                     # :299
        self.end_MLmodel = time.time()                               # :301
        if opt["loss_type"] == "pde_loss" and not return_states:
            return self._pde_tail(data, x_phys)                      # :307-342
        if return_states:
            return x_phys, states, edge_index
        return x_phys

    def _pde_tail(self, data, x_phys):
        """`loss_type == 'pde_loss'` (GNN.py:307-342): per mesh, the differentiable FEM solve on the relocated nodes
        (1-D: difFEM_1d.py:159-186 = fem1d_oracle.torch_fem_1d; 2-D: difFEM_2d.py:345-372 = fem2d_oracle.torch_fem_2d,
        the line-by-line restatement, or with opt['oracle_fem2d'] == 'fast' the vectorised formulation of
        fem2d_fast.py), the 2-D solution re-ordered to the fine mesh's node order (`reshape_grid_to_fd_tensor` of a
        [Q*Q, 1] column, utils_data.py:143-159 = a gather by argsort(mapping_tensor_fine)); returns
        (coeffs, x_phys, sol) concatenated over the batch."""
        import numpy as np
        from . import fem1d_oracle, fem2d_fast, fem2d_oracle
        opt, dim = self.opt, self.dim
        Q = opt["eval_quad_points"]
        coefs, xs, sols = [], [], []
        if dim == 2:
            x0 = torch.linspace(0, 1, Q)
            X, Y = torch.meshgrid(x0, x0, indexing="ij")                 # :185-188
            mesh = self.dataset.mesh
            cells = np.asarray(mesh.coordinates.cell_node_map().values)
            bc = np.asarray(mesh.bc_nodes)
            _, order = torch.sort(self.dataset.mapping_tensor_fine)
        for b in range(int(data.batch.max().item()) + 1):
            c_list = [torch.as_tensor(np.asarray(c)) for c in data.pde_params["centers"][b]]   # :311-316
            s_list = [torch.as_tensor(np.asarray(s)) for s in data.pde_params["scales"][b]]
            xb = x_phys.squeeze()[data.batch == b]                   # :320,327
            if dim == 1:
                coef, sol, _, _ = fem1d_oracle.torch_fem_1d(xb, torch.linspace(0, 1, Q), c_list, s_list,
                                                            load_quad_points=opt["load_quad_points"],
                                                            stiff_quad_points=opt["stiff_quad_points"])
            elif opt.get("oracle_fem2d", "reference") == "fast":
                coef, sol = fem2d_fast.fem2d_fast(cells, bc, xb, [X, Y], opt["load_quad_points"],
                                                  torch.stack(c_list).float(), torch.stack(s_list).float())
                sol = sol.reshape(-1)[order]
            else:
                coef, sol = fem2d_oracle.torch_fem_2d(cells, bc, xb, [X, Y], opt["load_quad_points"], c_list, s_list)
                sol = sol.reshape(-1)[order]                         # :333
            coefs.append(coef)
            xs.append(xb)
            sols.append(sol)
        return torch.cat(coefs, dim=0), torch.cat(xs, dim=0), torch.cat(sols, dim=0)


# --------------------------------------------------------------------------------------
# cross-checks
# --------------------------------------------------------------------------------------
def dense_layer(x: Tensor, edge_index: Tensor, Wq: Tensor, bq: Tensor, Wk: Tensor, bk: Tensor,
                inv_temp: float = 1.0) -> Tuple[Tensor, Tensor]:
    """Second, independent formulation of one layer: dense masked softmax.
    Returns (res = A x - x, A) with A[i, j] = sum of alpha over the (possibly repeated) edges j->i."""
    N, C = x.shape
    q = x @ Wq.T + bq
    k = x @ Wk.T + bk
    S = (q @ k.T) / math.sqrt(Wq.shape[0]) * inv_temp            # S[i, j] = <q_i, k_j>/sqrt(C)/T
    mult = torch.zeros(N, N, dtype=x.dtype)
    mult.index_put_((edge_index[1], edge_index[0]), torch.ones(edge_index.shape[1], dtype=x.dtype),
                    accumulate=True)                              # multiplicity of edge j->i
    neg = torch.full_like(S, -float("inf"))
    Sm = torch.where(mult > 0, S, neg)
    m = Sm.max(dim=1, keepdim=True).values
    m = torch.where(torch.isfinite(m), m, torch.zeros_like(m))
    P = torch.exp(Sm - m) * mult
    Z = P.sum(dim=1, keepdim=True) + 1e-16
    A = P / Z
    return A @ x - x, A


def csr_by_destination(edge_index: Tensor, num_nodes: int):
    """Oracle for the graph builder: stable sort of the edge list on the destination.
    Returns (rowptr int32 [N+1], col int32 [E], eid int32 [E]) -- `eid[s]` is the position in
    `edge_index` of the edge stored in CSR slot s.  Because `scatter_add_` on CPU accumulates
    each destination row in edge-list order, this is also the reference's summation order."""
    dst = edge_index[1]
    eid = torch.sort(dst, stable=True).indices
    col = edge_index[0][eid]
    counts = torch.bincount(dst, minlength=num_nodes)
    rowptr = torch.zeros(num_nodes + 1, dtype=torch.int64)
    rowptr[1:] = torch.cumsum(counts, 0)
    return rowptr.to(torch.int32), col.to(torch.int32), eid.to(torch.int32)


def csc_by_source(edge_index: Tensor, num_nodes: int):
    """Transpose structure for the backward: stable sort on the source.
    Returns (t_rowptr int32 [N+1], t_dst int32 [E], t_eid int32 [E])."""
    src = edge_index[0]
    eid = torch.sort(src, stable=True).indices
    dst = edge_index[1][eid]
    counts = torch.bincount(src, minlength=num_nodes)
    rowptr = torch.zeros(num_nodes + 1, dtype=torch.int64)
    rowptr[1:] = torch.cumsum(counts, 0)
    return rowptr.to(torch.int32), dst.to(torch.int32), eid.to(torch.int32)


def mesh_loss(out: Tensor, target: Tensor, loss_fn: str = "l1") -> Tensor:
    """`run_GNN.py:80-84,103-106`."""
    if target.dim() == 1:
        target = target.unsqueeze(-1)
    return F.l1_loss(out, target) if loss_fn == "l1" else F.mse_loss(out, target)
