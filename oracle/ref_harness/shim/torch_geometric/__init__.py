"""TEST INFRASTRUCTURE -- minimal stand-in for torch-geometric 2.4.0 (pinned by the reference,
README.md:25; absent from this image).  Only what `src/GRAND_plus.py` and `src/GNN.py` import,
built on the primitive restatements in `oracle/pyg_semantics.py`.  Exists so that the
reference's OWN source files can be executed here to mint `tests/golden/` fixtures."""
__version__ = "2.4.0-shim"
