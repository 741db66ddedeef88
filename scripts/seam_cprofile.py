"""Where the host time of an inference-mode `model(data)` call goes (cProfile over many calls on a resident batch)."""
import cProfile
import os
import pstats
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from g_adaptivity_b200 import GNN, synth  # noqa: E402

md, B, burgers = ((200,), 4096, True) if "--cfg3" in sys.argv else ((15, 15), 1, False)
opt = synth.burgers_opt(md) if burgers else synth.default_opt(md)
opt.update(device="cuda:0", gad_store_alpha=False, gad_sync_timestamp=False)
ds = synth.SyntheticDataset(len(md), md)
torch.manual_seed(42)
model = GNN(ds, opt).to("cuda:0").eval()
data = synth.make_batch(md, B, seed=0, burgers=burgers).to("cuda:0")
with torch.no_grad():
    for _ in range(20):
        model(data)
    torch.cuda.synchronize()
    pr = cProfile.Profile()
    pr.enable()
    for _ in range(2000):
        model(data)
    pr.disable()
    torch.cuda.synchronize()
st = pstats.Stats(pr)
st.sort_stats("cumulative").print_stats(28)
