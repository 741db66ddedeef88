#!/usr/bin/env python
"""Forward-only dispatch sweep: cluster-resident kernel against the streaming wide-row chain, per mesh
size, batch and integrator (module seam, CUDA-graph replay through InferenceSession).
python scripts/fwd_sweep.py"""
import os, sys, json
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from g_adaptivity_b200 import GNN, synth


def run(md, batch, method, layers, no_cluster):
    opt = synth.default_opt(md, device="cuda:0", gad_store_alpha=False, ode_method=method, num_layers=layers,
                            gad_no_cluster=no_cluster)
    ds = synth.SyntheticDataset(2, md)
    torch.manual_seed(42)
    model = GNN(ds, opt).to("cuda:0").eval()
    data = synth.make_batch(md, batch, seed=0).to("cuda:0")
    with torch.no_grad():
        sess = model.inference_session(data)
        for _ in range(3):
            sess()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(10):
            sess()
        e1.record()
        torch.cuda.synchronize()
    return e0.elapsed_time(e1) / 10, getattr(sess.graph, "clf_C", None)


cases = [((64, 64), b) for b in (1, 4, 16, 64)] + [((100, 100), b) for b in (1, 2, 4, 8, 16, 64)] + \
        [((200, 200), b) for b in (1, 4)]
for method, layers in (("rk4", 16), ("euler", 4)):
    for md, b in cases:
        row = {"mesh": md, "batch": b, "method": method, "layers": layers}
        for nc in (False, True):
            try:
                ms, C = run(md, b, method, layers, nc)
                row["stream_ms" if nc else "cluster_ms"] = round(ms, 4)
                if not nc:
                    row["C"] = C
            except Exception as ex:  # noqa
                row["stream_err" if nc else "cluster_err"] = str(ex)[:80]
        print(json.dumps(row), flush=True)
