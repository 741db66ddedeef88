"""TEST INFRASTRUCTURE -- CPU restatement of the torch_geometric 2.4.0 primitives the
reference's hot path calls.  Never imported by the product package.

The arithmetic of `src/GRAND_plus.py:233,333` and `src/GNN.py:222-223` lives in a third-party
dependency that is absent from /root/reference and from this image:

    torch-geometric == 2.4.0   (pinned by the reference, README.md:25)

so its published algorithm is restated here, function by function:

* `torch_geometric.utils.scatter`  (utils/scatter.py): 'sum' -> `new_zeros(size).scatter_add_`,
  'max' -> `new_zeros(size).scatter_reduce_(..., 'amax', include_self=False)`;
* `torch_geometric.utils.softmax`  (utils/softmax.py): max over the *detached* source,
  `exp(src - max[index])`, `sum + 1e-16`, divide;
* `torch_geometric.utils.remove_self_loops / add_self_loops` (utils/loop.py): mask keeps order,
  `arange(N)` loops appended after the existing edges;
* `MessagePassing.propagate` with `flow='source_to_target'`, `node_dim=0`, `aggr='add'`
  (nn/conv/message_passing.py): `x_j = x[edge_index[0]]`, `x_i = x[edge_index[1]]`,
  `index = edge_index[1]`, aggregate = scatter-sum over `index` with `dim_size = N`.

Parity status: these restatements cannot be checked against PyG offline ("from knowledge of
PyG 2.4.0"); they are cross-checked against a dense-matrix formulation in
`oracle/gnn_oracle.py::dense_layer` and exercised by the reference's own `GRAND_plus.py` /
`GNN.py` source through `oracle/ref_harness/`.
"""
from __future__ import annotations

from typing import Optional, Tuple

import torch
from torch import Tensor


def _broadcast(index: Tensor, ref: Tensor, dim: int) -> Tensor:
    size = [1] * ref.dim()
    size[dim] = -1
    return index.view(size).expand_as(ref)


def scatter(src: Tensor, index: Tensor, dim: int = 0, dim_size: Optional[int] = None,
            reduce: str = "sum") -> Tensor:
    """torch_geometric.utils.scatter (2.4.0), 'sum'/'add'/'max'/'mean' only."""
    dim = src.dim() + dim if dim < 0 else dim
    if dim_size is None:
        dim_size = int(index.max()) + 1 if index.numel() > 0 else 0
    size = list(src.size())
    size[dim] = dim_size
    if reduce in ("sum", "add"):
        return src.new_zeros(size).scatter_add_(dim, _broadcast(index, src, dim), src)
    if reduce == "mean":
        count = src.new_zeros(dim_size).scatter_add_(0, index, src.new_ones(src.size(dim)))
        count = count.clamp(min=1)
        out = src.new_zeros(size).scatter_add_(dim, _broadcast(index, src, dim), src)
        return out / _broadcast(count, out, dim)
    if reduce in ("max", "amax"):
        return src.new_zeros(size).scatter_reduce_(dim, _broadcast(index, src, dim), src,
                                                   reduce="amax", include_self=False)
    raise ValueError(reduce)


def softmax(src: Tensor, index: Optional[Tensor] = None, ptr: Optional[Tensor] = None,
            num_nodes: Optional[int] = None, dim: int = 0) -> Tensor:
    """torch_geometric.utils.softmax (2.4.0), `index` form (ptr is None on this path)."""
    assert ptr is None and index is not None
    N = num_nodes if num_nodes is not None else (int(index.max()) + 1 if index.numel() else 0)
    src_max = scatter(src.detach(), index, dim, dim_size=N, reduce="max")
    out = src - src_max.index_select(dim, index)
    out = out.exp()
    out_sum = scatter(out, index, dim, dim_size=N, reduce="sum") + 1e-16
    out_sum = out_sum.index_select(dim, index)
    return out / out_sum


def remove_self_loops(edge_index: Tensor, edge_attr: Optional[Tensor] = None) -> Tuple[Tensor, Optional[Tensor]]:
    mask = edge_index[0] != edge_index[1]
    edge_index = edge_index[:, mask]
    return edge_index, (None if edge_attr is None else edge_attr[mask])


def add_self_loops(edge_index: Tensor, edge_attr: Optional[Tensor] = None, fill_value=None,
                   num_nodes: Optional[int] = None) -> Tuple[Tensor, Optional[Tensor]]:
    assert edge_attr is None
    N = num_nodes if num_nodes is not None else int(edge_index.max()) + 1
    loop_index = torch.arange(0, N, dtype=torch.long, device=edge_index.device)
    loop_index = loop_index.unsqueeze(0).repeat(2, 1)
    return torch.cat([edge_index, loop_index], dim=1), None


def propagate_add(edge_index: Tensor, message_fn, num_nodes: int, **node_kwargs) -> Tensor:
    """`MessagePassing.propagate` for flow='source_to_target', node_dim=0, aggr='add'.

    `node_kwargs` are `[N, ...]` tensors; the message function receives `<name>_i`
    (gathered with `edge_index[1]`) and `<name>_j` (gathered with `edge_index[0]`)."""
    j, i = edge_index[0], edge_index[1]
    coll = {}
    for name, t in node_kwargs.items():
        coll[name + "_j"] = t.index_select(0, j)
        coll[name + "_i"] = t.index_select(0, i)
    msg = message_fn(index=i, size_i=num_nodes, **coll)
    return scatter(msg, i, dim=0, dim_size=num_nodes, reduce="sum")
