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


cases = [((64, 64), b) for b in (1, 4, 16)] + [((100, 100), b) for b in (1, 2, 4, 8, 16)] + \
        [((200, 200), b) for b in (1, 2)]
# routes: the cluster kernel (forced), the persistent one-launch streaming kernel, the chain of dependent launches, and
# what graph.stream_fwd_preferred picks when left alone
ROUTES = (("cluster_ms", {"GAD_FWD_POLICY": "cluster"}, False), ("persist_ms", {}, True),
          ("chain_ms", {"GAD_WIDE_PERSIST": "0"}, True), ("default_ms", {}, False))
for method, layers in (("rk4", 16), ("euler", 4)):
    for md, b in cases:
        row = {"mesh": md, "batch": b, "method": method, "layers": layers}
        for key, env, nc in ROUTES:
            for k in ("GAD_FWD_POLICY", "GAD_WIDE_PERSIST"):
                os.environ.pop(k, None)
            os.environ.update(env)
            try:
                ms, C = run(md, b, method, layers, nc)
                row[key] = round(ms, 4)
                if key == "cluster_ms":
                    row["C"] = C
            except Exception as ex:  # noqa
                row[key.replace("_ms", "_err")] = str(ex)[:80]
        best = min((row[k], k) for k in ("cluster_ms", "persist_ms", "chain_ms") if k in row)
        row["policy_loss_pct"] = round(100.0 * (row.get("default_ms", best[0]) / best[0] - 1.0), 1)
        print(json.dumps(row), flush=True)
