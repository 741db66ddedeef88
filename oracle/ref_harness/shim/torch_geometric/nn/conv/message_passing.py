"""torch_geometric.nn.conv.MessagePassing (2.4.0), the slice the reference exercises:
flow='source_to_target', dense `edge_index`, `aggr='add'`, no hooks, no fused path.

propagate():  size <- N;  for every argument `a_i` / `a_j` of `message`, gather
`kwargs['a']` along `node_dim` with `edge_index[1]` / `edge_index[0]`;  special arguments
`edge_index`, `index` (= edge_index[1]), `ptr` (None), `size_i`/`dim_size` (= N);  other names
are passed through;  aggregate = scatter-sum over `index`;  update = identity."""
import inspect

import torch.nn as nn

from oracle.pyg_semantics import scatter


class MessagePassing(nn.Module):
    def __init__(self, aggr="add", flow="source_to_target", node_dim=-2, **kwargs):
        # 2.4.0 accepts aggr_kwargs / decomposed_layers explicitly; the reference additionally
        # leaks `global_feat_dim=` into **kwargs (src/GNN.py:118) -- accepted and ignored here.
        super().__init__()
        assert aggr == "add" and flow == "source_to_target"
        self.aggr, self.flow, self.node_dim = aggr, flow, node_dim
        self._msg_params = [p for p in inspect.signature(self.message).parameters]

    def reset_parameters(self):
        pass

    def propagate(self, edge_index, size=None, **kwargs):
        j, i = edge_index[0], edge_index[1]
        N = None
        for v in kwargs.values():
            if hasattr(v, "size") and v is not None:
                N = v.size(self.node_dim)
                break
        coll = {}
        for name in self._msg_params:
            if name.endswith("_i") and name[:-2] in kwargs:
                src = kwargs[name[:-2]]
                coll[name] = None if src is None else src.index_select(self.node_dim, i)
            elif name.endswith("_j") and name[:-2] in kwargs:
                src = kwargs[name[:-2]]
                coll[name] = None if src is None else src.index_select(self.node_dim, j)
            elif name == "edge_index":
                coll[name] = edge_index
            elif name == "index":
                coll[name] = i
            elif name == "ptr":
                coll[name] = None
            elif name in ("size_i", "dim_size"):
                coll[name] = N
            elif name == "size_j":
                coll[name] = N
            elif name in kwargs:
                coll[name] = kwargs[name]
        out = self.message(**coll)
        out = scatter(out, i, dim=self.node_dim, dim_size=N, reduce="sum")
        return self.update(out)

    def message(self, x_j):
        return x_j

    def update(self, inputs):
        return inputs
