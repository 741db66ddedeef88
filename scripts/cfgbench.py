#!/usr/bin/env python
"""Throughput of the BASELINE.json configs that are NOT the bench line (parity-test cases), for
DESIGN.md section 6: module-seam forward (cfg 1, 3, 4) and trainer step (cfg 2, 5 slice), CUDA
events on the launching stream, inputs device-resident.   python scripts/cfgbench.py [--quick]"""
import argparse
import json
import os
import statistics
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch

PEAK = 6539.9


def fwd_bytes(N, E, ce, fevals, in_dim, dim):
    return fevals * (8 * ce + 4 * E / N + 4) + 4 * in_dim + 4 * dim


def time_forward(md, B, over, burgers, iters):
    from g_adaptivity_b200 import GNN, synth
    dev = torch.device("cuda", 0)
    opt = synth.burgers_opt(md) if burgers else synth.default_opt(md)
    opt.update(device="cuda:0", gad_store_alpha=False, **over)
    ds = synth.SyntheticDataset(len(md), md)
    torch.manual_seed(42)
    model = GNN(ds, opt).to(dev).eval()
    data = synth.make_batch(md, B, seed=0, burgers=burgers).to(dev)
    with torch.no_grad():
        for _ in range(3):
            model(data)
        torch.cuda.synchronize()
        ts = []
        for _ in range(iters):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            model(data)
            e1.record()
            torch.cuda.synchronize()
            ts.append(e0.elapsed_time(e1))
        sess = model.inference_session(data)
        for _ in range(3):
            sess()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(50):
            sess(uu=data.uu_tensor)
        e1.record()
        torch.cuda.synchronize()
        ms_sess = e0.elapsed_time(e1) / 50
    g = model.last_graph
    ms = statistics.median(ts)
    fevals = opt["num_layers"] * (4 if opt.get("ode_method", "euler") == "rk4" else 1)
    bpn = fwd_bytes(g.N, g.E, model.live, fevals, sum(model.in_dims), model.dim)
    return {"nodes": g.N, "edges": g.E, "ms_per_call": round(ms, 4), "gnodes_per_s": round(g.N / ms / 1e6, 4),
            "bytes_per_node": round(bpn, 1), "roofline_frac": round(g.N * bpn / (ms * 1e-3) / 1e9 / PEAK, 4),
            "tiles": g.T if g.tile_ptr is not None else 0, "ms_per_call_session": round(ms_sess, 4),
            "gnodes_per_s_session": round(g.N / ms_sess / 1e6, 4),
            "roofline_frac_session": round(g.N * bpn / (ms_sess * 1e-3) / 1e9 / PEAK, 4)}


def time_train(md, B, ring, iters):
    from g_adaptivity_b200 import GNN, synth
    from g_adaptivity_b200.trainer import DeformerTrainer
    dev = torch.device("cuda", 0)
    opt = synth.default_opt(md)
    opt.update(device="cuda:0", gad_store_alpha=False)
    ds = synth.SyntheticDataset(len(md), md)
    torch.manual_seed(42)
    model = GNN(ds, opt).to(dev)
    tr = DeformerTrainer(model)
    sids = [tr.add_batch(synth.make_batch(md, B, seed=100 + r)) for r in range(ring)]
    tr.capture_epoch(sids)
    for _ in range(3):
        tr.run_epoch(sids)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(tr.stream)
    for _ in range(iters):
        tr.run_epoch(sids)
    e1.record(tr.stream)
    tr.synchronize()
    ms = e0.elapsed_time(e1) / (iters * ring)
    s = tr.slots[0]
    N, E = s.N, s.graph.E
    d = E / N
    bpn = 4 * ((8 * 4 + 4 * d + 4) + (12 * 4 + 2 * (4 * d + 4))) + 4 * 4 + 8 * 2
    return {"nodes": N, "edges": E, "ms_per_step": round(ms, 4), "gnodes_per_s": round(N / ms / 1e6, 4),
            "bytes_per_node": round(bpn, 1), "roofline_frac": round(N * bpn / (ms * 1e-3) / 1e9 / PEAK, 4),
            "tiles": s.graph.T}


def time_module_train_fresh(md, B, iters, content_cache, shared_topology=False):
    """The reference's loop shape: a FRESH Batch object per iteration, model(data) -> L1 loss ->
    backward -> torch.optim.Adam.step(), all through the module seam (host tensors in)."""
    import torch.nn.functional as F
    from g_adaptivity_b200 import GNN, synth
    dev = torch.device("cuda", 0)
    opt = synth.default_opt(md)
    opt.update(device="cuda:0", gad_store_alpha=False, gad_content_cache=content_cache, gad_shared_topology=shared_topology)
    ds = synth.SyntheticDataset(len(md), md)
    torch.manual_seed(42)
    model = GNN(ds, opt).to(dev).train()
    optim = torch.optim.Adam(model.parameters(), lr=1e-3)
    base = synth.make_batch(md, B, seed=0)
    base.pin_memory()

    def fresh():
        b = base.clone()
        b.pin_memory()
        return b

    batches = [fresh() for _ in range(iters + 3)]
    ts = []
    for it, data in enumerate(batches):
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        optim.zero_grad(set_to_none=True)
        out = model(data)
        F.l1_loss(out, data.x_phys.to(dev, non_blocking=True)).backward()
        optim.step()
        torch.cuda.synchronize()
        if it >= 3:
            ts.append(time.perf_counter() - t0)
    ms = 1e3 * statistics.median(ts)
    N = base.x_comp.shape[0]
    return {"nodes": N, "ms_per_step": round(ms, 4), "gnodes_per_s": round(N / ms / 1e6, 4),
            "graph_builds": model._graphs.misses, "content_hits": getattr(model._graphs, "misses_identity", 0),
            "shared_hits": getattr(model._graphs, "shared_hits", 0)}


def time_fem1d(B, n, G=1, K=101, Q=101, iters=10):
    """Next row f1 (1-D): batched differentiable FEM solve behind pde_loss, forward + backward (the CPU
    restatement of torch_FEM_1D the reference loops over is timed by tests/tools/fem1d_cpu_baseline.py)."""
    import numpy as np
    import torch.nn.functional as F
    from g_adaptivity_b200 import fem1d
    rng = np.random.default_rng(0)
    xs = np.tile(np.linspace(0, 1, n), (B, 1))
    xs[:, 1:-1] += (rng.random((B, n - 2)) - 0.5) * 0.3 / (n - 1)
    centers = torch.from_numpy(rng.uniform(0.3, 0.7, (B, G)).astype(np.float32)).cuda()
    scales = torch.from_numpy(rng.uniform(0.05, 0.2, (B, G)).astype(np.float32)).cuda()
    x = torch.from_numpy(xs.astype(np.float32)).reshape(-1).cuda().requires_grad_(True)
    quad = torch.linspace(0, 1, Q).cuda()
    tgt = torch.zeros(B * Q, device="cuda")

    def step():
        x.grad = None
        _, sol = fem1d.fem1d_solve(x, centers, scales, quad, n, K)
        F.mse_loss(sol, tgt).backward()

    for _ in range(3):
        step()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        step()
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / iters
    return {"meshes": B, "nodes_per_mesh": n, "gpu_ms_fwd_bwd": round(ms, 4), "meshes_per_s": round(B / ms * 1e3, 1),
            "cpu_baseline": "tests/tools/fem1d_cpu_baseline.py (6.5 ms per mesh on the box's host)"}


def time_pde_loss_step(md, B, iters):
    """loss_type='pde_loss' on 1-D meshes through the module seam: model(data) (deformer + batched FEM) ->
    mse(sol, u_true_fine) -> backward -> torch Adam; inputs device-resident."""
    import torch.nn.functional as F
    from g_adaptivity_b200 import GNN, synth
    dev = torch.device("cuda", 0)
    opt = synth.default_opt(md)
    opt.update(device="cuda:0", gad_store_alpha=False, loss_type="pde_loss")
    ds = synth.SyntheticDataset(len(md), md)
    torch.manual_seed(42)
    model = GNN(ds, opt).to(dev).train()
    optim = torch.optim.Adam(model.parameters(), lr=1e-3)
    data = synth.make_batch(md, B, seed=0).to(dev)
    tgt = data.u_true_fine_tensor
    ts = []
    for it in range(iters + 3):
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        optim.zero_grad(set_to_none=True)
        coeffs, xp, sol = model(data)
        F.mse_loss(sol, tgt).backward()
        optim.step()
        torch.cuda.synchronize()
        if it >= 3:
            ts.append(time.perf_counter() - t0)
    ms = 1e3 * statistics.median(ts)
    return {"meshes": B, "nodes": B * md[0], "ms_per_step": round(ms, 3), "meshes_per_s": round(B / ms * 1e3, 1)}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--quick", action="store_true")
    a = ap.parse_args()
    it = 5 if a.quick else 20
    res = {}
    res["cfg1_15x15_fwd"] = time_forward((15, 15), 1, {}, False, it)
    res["cfg3_burgers200_b4096_fwd"] = time_forward((200,), 4096, {}, True, it)
    res["cfg4_200x200_rk4x64_fwd"] = time_forward((200, 200), 1, {"ode_method": "rk4", "num_layers": 64}, False, it)
    res["cfg2_30x30_b256_train"] = time_train((30, 30), 256, 8, it)
    res["cfg5_50x50_b1024_train"] = time_train((50, 50), 1024, 2, max(2, it // 4))
    res["fem1d_pde_loss_b4096_n200"] = time_fem1d(4096, 200)
    res["pde_loss_train_step_module_seam_b4096_n200"] = time_pde_loss_step((200,), 4096, it)
    res["cfg2_module_seam_train_fresh_batches_content_cache"] = time_module_train_fresh((30, 30), 256, it, True)
    res["cfg2_module_seam_train_fresh_batches_shared_topology"] = time_module_train_fresh((30, 30), 256, it, True, True)
    res["cfg2_module_seam_train_fresh_batches_identity_cache_only"] = time_module_train_fresh((30, 30), 256, it, False)
    for k, v in res.items():
        print(k, json.dumps(v), flush=True)


if __name__ == "__main__":
    main()
