"""Stand-in for torch_scatter (only used by the dead `G2` class, `src/GRAND_plus.py:30-31`)."""
from oracle.pyg_semantics import scatter as _scatter


def scatter(src, index, dim=-1, out=None, dim_size=None, reduce="sum"):
    assert out is None
    return _scatter(src, index, dim, dim_size, reduce)
