"""TEST INFRASTRUCTURE: runs the reference's own GRAND_plus.py / GNN.py (read from
/root/reference, never copied) on top of a torch-geometric 2.4.0 shim, in THIS container only,
to mint the fixtures under tests/golden/."""
