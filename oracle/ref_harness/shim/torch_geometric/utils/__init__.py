from oracle.pyg_semantics import softmax, add_self_loops, remove_self_loops, scatter  # noqa: F401
