"""torch_geometric.nn.conv.TransformerConv (2.4.0), restated for the configuration
`GRAND_conv` uses (`src/GRAND_plus.py:371-372`): heads=1, concat=False, beta=False, dropout=0,
edge_dim=None, bias=False, root_weight=False."""
import math

import torch.nn.functional as F
from torch import Tensor

from ..dense.linear import Linear
from .message_passing import MessagePassing
from oracle.pyg_semantics import softmax


class TransformerConv(MessagePassing):
    def __init__(self, in_channels, out_channels, heads=1, concat=True, beta=False, dropout=0.0,
                 edge_dim=None, bias=True, root_weight=True, **kwargs):
        kwargs.setdefault("aggr", "add")
        super().__init__(node_dim=0, **kwargs)
        self.in_channels, self.out_channels, self.heads = in_channels, out_channels, heads
        self.beta = beta and root_weight
        self.root_weight, self.concat, self.dropout, self.edge_dim = root_weight, concat, dropout, edge_dim
        self._alpha = None
        if isinstance(in_channels, int):
            in_channels = (in_channels, in_channels)
        self.lin_key = Linear(in_channels[0], heads * out_channels)
        self.lin_query = Linear(in_channels[1], heads * out_channels)
        self.lin_value = Linear(in_channels[0], heads * out_channels)
        assert edge_dim is None
        self.lin_edge = self.register_parameter("lin_edge", None)
        if concat:
            self.lin_skip = Linear(in_channels[1], heads * out_channels, bias=bias)
        else:
            self.lin_skip = Linear(in_channels[1], out_channels, bias=bias)
        assert not self.beta
        self.lin_beta = self.register_parameter("lin_beta", None)
        self.reset_parameters()

    def reset_parameters(self):
        super().reset_parameters()
        self.lin_key.reset_parameters()
        self.lin_query.reset_parameters()
        self.lin_value.reset_parameters()
        self.lin_skip.reset_parameters()

    def forward(self, x, edge_index, edge_attr=None, return_attention_weights=None):
        H, C = self.heads, self.out_channels
        if isinstance(x, Tensor):
            x = (x, x)
        query = self.lin_query(x[1]).view(-1, H, C)
        key = self.lin_key(x[0]).view(-1, H, C)
        value = self.lin_value(x[0]).view(-1, H, C)
        out = self.propagate(edge_index, query=query, key=key, value=value, edge_attr=edge_attr, size=None)
        alpha = self._alpha
        self._alpha = None
        if self.concat:
            out = out.view(-1, self.heads * self.out_channels)
        else:
            out = out.mean(dim=1)
        if self.root_weight:
            out = out + self.lin_skip(x[1])
        if isinstance(return_attention_weights, bool):
            assert alpha is not None
            return out, (edge_index, alpha)
        return out

    def message(self, query_i, key_j, value_j, edge_attr, index, ptr, size_i):
        alpha = (query_i * key_j).sum(dim=-1) / math.sqrt(self.out_channels)
        alpha = softmax(alpha, index, ptr, size_i)
        self._alpha = alpha
        alpha = F.dropout(alpha, p=self.dropout, training=self.training)
        out = value_j
        out = out * alpha.view(-1, self.heads, 1)
        return out
