import ctypes as C, sys
sys.path.insert(0, '/root/repo')
import torch
torch.cuda.init()
from g_adaptivity_b200 import _lib
lib = _lib.load()
for (Cs, S, T) in [(2, 1800, 480), (2, 3168, 512), (4, 2500, 512), (4, 1250, 320), (8, 1250, 320), (8, 1250, 512), (8, 2500, 512), (16, 2500, 512), (16, 640, 192), (16, 1250, 320)]:
    n = C.c_int(0)
    rc = lib.gad_cluster_occupancy(4, Cs, S, T, C.byref(n))
    print(f"C={Cs:2d} slab={S:5d} threads={T}: rc={rc} max active clusters={n.value} -> CTAs {n.value*Cs}")
