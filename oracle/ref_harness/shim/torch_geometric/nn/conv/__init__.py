from .message_passing import MessagePassing  # noqa: F401
from .transformer_conv import TransformerConv  # noqa: F401


class _OffPath:
    """Convs the hot path never builds (`conv_type` in GCN/GAT/GAT_plus are out of scope)."""

    def __init__(self, *a, **k):
        raise NotImplementedError(f"{type(self).__name__} is not part of the deformer hot path")


class GATConv(_OffPath):
    pass


class GCNConv(_OffPath):
    pass
