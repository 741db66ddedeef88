"""Import the reference's real `GRAND_plus` and `GNN` modules from /root/reference/src.

They cannot be imported as they stand: their module tops pull in torch_geometric,
torch_scatter, Firedrake, torchquad, wandb, matplotlib.  This loader
  * puts `oracle/ref_harness/shim` (torch_geometric / torch_scatter stand-ins built on
    `oracle/pyg_semantics.py`) in front of `sys.path`;
  * registers empty stand-ins for the modules that are imported but never reached on the hot
    path: `firedrake_difFEM.difFEM_1d/_2d` (pde_loss tail), `wandb`, `matplotlib`; `feature_extractors`
    (global CNN) and `utils_data` (grid reshapes for the CNN) are the reference's real modules;
  * imports `params`, `GRAND_plus`, `GNN` from the read-only reference tree.
Nothing is copied: the reference source is executed where it lies.  Only usable where
/root/reference exists (this container, not the GPU box)."""
from __future__ import annotations

import contextlib
import io
import os
import sys
import types

REFERENCE_SRC = "/root/reference/src"
_SHIM = os.path.join(os.path.dirname(os.path.abspath(__file__)), "shim")
_REPO = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def available() -> bool:
    return os.path.isfile(os.path.join(REFERENCE_SRC, "GNN.py"))


def _stub(name: str, **attrs) -> types.ModuleType:
    m = types.ModuleType(name)
    m.__dict__.update(attrs)
    sys.modules[name] = m
    return m


def _unreachable(name):
    def f(*a, **k):
        raise RuntimeError(f"{name} is off the deformer hot path and is not available in the harness")
    f.__name__ = name
    return f


def load():
    """Returns (GNN_module, GRAND_plus_module, params_module) of the reference."""
    if not available():
        raise RuntimeError("/root/reference is not present: the harness only runs in the build container")
    for p in (_REPO, _SHIM, REFERENCE_SRC):
        if p not in sys.path:
            sys.path.insert(0, p)
    # shim must win over anything else named torch_geometric
    sys.path.remove(_SHIM)
    sys.path.insert(0, _SHIM)
    pkg = _stub("firedrake_difFEM")
    pkg.__path__ = []
    _stub("firedrake_difFEM.difFEM_1d", torch_FEM_1D=_unreachable("torch_FEM_1D"))
    _stub("firedrake_difFEM.difFEM_2d", torch_FEM_2D=_unreachable("torch_FEM_2D"))
    # feature_extractors.py (global CNN, row f3) and utils_data.py (grid reorderings) are the reference's REAL
    # modules: plain torch once wandb / matplotlib (imported at the top of utils_data.py, unused on this path) answer
    for name in ("wandb", "matplotlib", "matplotlib.pyplot"):
        if name not in sys.modules:
            _stub(name)
    sys.modules["matplotlib"].pyplot = sys.modules["matplotlib.pyplot"]
    import importlib
    for name in ("feature_extractors", "utils_data"):
        sys.modules.pop(name, None)
    params = importlib.import_module("params")
    fe = importlib.import_module("feature_extractors")
    ud = importlib.import_module("utils_data")
    assert os.path.realpath(fe.__file__).startswith("/root/reference/"), fe.__file__
    assert os.path.realpath(ud.__file__).startswith("/root/reference/"), ud.__file__
    grand = importlib.import_module("GRAND_plus")
    gnn = importlib.import_module("GNN")
    assert os.path.realpath(gnn.__file__).startswith("/root/reference/"), gnn.__file__
    assert os.path.realpath(grand.__file__).startswith("/root/reference/"), grand.__file__
    return gnn, grand, params


@contextlib.contextmanager
def quiet():
    """`get_arg_list` prints its argument on every model construction (params.py:191)."""
    with contextlib.redirect_stdout(io.StringIO()):
        yield
