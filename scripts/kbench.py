#!/usr/bin/env python
"""Kernel micro-benchmark: CUDA-event time of the fused forward / backward kernels (and the whole
graph-replayed train step) for one workload.  Used to compare kernel variants on the GPU box:

    GAD_LIB=/path/to/libvariant.so python scripts/kbench.py --mesh 30 30 --batch 256

(Parity against the CPU oracle lives in tests/; this script only measures.)"""
import argparse
import copy
import json
import os
import statistics
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import torch


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--mesh", type=int, nargs="+", default=[30, 30])
    ap.add_argument("--batch", type=int, default=256)
    ap.add_argument("--layers", type=int, default=4)
    ap.add_argument("--ring", type=int, default=8)
    ap.add_argument("--iters", type=int, default=200)
    ap.add_argument("--tile-nodes", type=int, default=None)
    ap.add_argument("--no-step", action="store_true", help="skip the graph-captured whole-step timing")
    ap.add_argument("--tag", default=os.environ.get("GAD_LIB", "default"))
    a = ap.parse_args()

    from g_adaptivity_b200 import GNN, _lib, synth
    from g_adaptivity_b200.trainer import DeformerTrainer

    md = tuple(a.mesh)
    dev = torch.device("cuda", 0)
    burgers = len(md) == 1
    opt = (synth.burgers_opt(md) if burgers else synth.default_opt(md))
    opt.update(num_layers=a.layers, device="cuda:0", gad_store_alpha=False, loss_type="mesh_loss")
    if a.tile_nodes:
        opt["gad_tile_nodes"] = a.tile_nodes
    ds = synth.SyntheticDataset(len(md), md)
    torch.manual_seed(42)
    model = GNN(ds, opt).to(dev)
    sd0 = copy.deepcopy(model.state_dict())
    tr = DeformerTrainer(model)
    batches = [synth.make_batch(md, a.batch, seed=100 + r, burgers=burgers) for r in range(a.ring)]
    for b in batches:
        tr.add_batch(b)
    lib, P = _lib.load(), _lib.ptr
    st = tr.stream.cuda_stream
    res = {"tag": a.tag, "mesh": list(md), "batch": a.batch, "nodes": tr.slots[0].N, "tiles": tr.slots[0].graph.T,
           "max_tile_nodes": tr.slots[0].graph.max_tile_nodes}

    from g_adaptivity_b200 import functional as GF
    ell = GF.use_ell(tr.slots[0].graph, tr.CE)
    res["ell"] = bool(ell)

    def fwd(s):
        g = s.graph
        if ell:
            lib.gad_deform_fwd_ell(P(g.ell_in), s.N, P(g.tile_ptr), g.T, g.max_tile_nodes, g.ell_deg, P(s.states),
                                   model.dim, tr.CE, P(tr.Mu), tr.Lw, P(tr.tau), tr.L, 0, P(s.x_phys), P(s.states), st)
            return
        lib.gad_deform_fwd(P(g.rowptr), P(g.col_walk), s.N, g.E, P(g.tile_ptr), g.T, g.max_tile_nodes, g.max_tile_edges,
                           P(s.states), model.dim, tr.CE, P(tr.Mu), tr.Lw, P(tr.tau), tr.L, 0, P(s.x_phys), P(s.states),
                           P(s.fwd_ws), s.fwd_ws_bytes, st)

    def train(s):
        g = s.graph
        cnt = s.N * model.dim
        lib.gad_deform_train_ell(P(g.ell_in), P(g.ell_out), s.N, P(g.tile_ptr), g.T, g.max_tile_nodes, g.ell_deg,
                                 P(s.x_comp), P(s.f), P(s.uu), None, None, P(s.target), model.dim, tr.CE, P(tr.Mu), tr.Lw,
                                 P(tr.tau), tr.L, 0, 1.0 / cnt, 1.0 / cnt, P(s.states), P(tr.gMu), P(tr.gtau),
                                 P(s.loss), None, P(s.bwd_ws), s.bwd_ws_bytes, st)

    def bwd(s):
        g = s.graph
        if ell:
            lib.gad_deform_bwd_ell(P(g.ell_in), P(g.ell_out), s.N, P(g.tile_ptr), g.T, g.max_tile_nodes, g.ell_deg,
                                   P(s.states), P(s.g_out), model.dim, tr.CE, P(tr.Mu), None, tr.Lw, P(tr.tau), tr.L,
                                   P(tr.gMu), P(tr.gtau), None, P(s.bwd_ws), s.bwd_ws_bytes, st)
            return
        lib.gad_deform_bwd(P(g.rowptr), P(g.col_walk), P(g.t_rowptr), P(g.t_dst_walk), s.N, g.E, P(g.tile_ptr), g.T,
                           g.max_tile_nodes, g.max_tile_edges, P(s.states), P(s.g_out), model.dim, tr.CE, P(tr.Mu),
                           tr.Lw, P(tr.tau), tr.L, P(tr.gMu), P(tr.gtau), None, P(s.bwd_ws), s.bwd_ws_bytes, st)

    with torch.cuda.stream(tr.stream):
        for s in tr.slots:                      # one full eager step per slot: states / g_out valid
            tr._issue(s, st, with_optimizer=False)
    tr.synchronize()
    if ell:   # g_out for the stand-alone backward timing (the fused train kernel never materialises it)
        with torch.cuda.stream(tr.stream):
            for s in tr.slots:
                lib.gad_mesh_loss(P(s.x_phys), P(s.target), s.N * model.dim, 0, 1.0 / (s.N * model.dim), P(s.loss),
                                  P(s.g_out), P(s.loss_ws), st)
        tr.synchronize()
    for name, fn in (("fwd_us", fwd), ("bwd_us", bwd)) + ((("train_us", train),) if ell else ()):
        evs = []
        with torch.cuda.stream(tr.stream):
            torch.cuda._sleep(int(2e7))     # let the host run ahead: event pairs then time the device only
            for i in range(a.iters):
                s = tr.slots[i % a.ring]
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record(tr.stream)
                fn(s)
                e1.record(tr.stream)
                evs.append((e0, e1))
        tr.synchronize()
        ts = [1e3 * e0.elapsed_time(e1) for e0, e1 in evs[10:]]
        res[name] = round(statistics.median(ts), 2)
        res[name.replace("_us", "_p10")] = round(sorted(ts)[len(ts) // 10], 2)
    # whole step, graph replay
    if a.no_step:
        print(json.dumps(res), flush=True)
        return
    for sid in range(a.ring):
        tr.capture(sid)
    tr.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    for i in range(20):
        tr.step(i % a.ring)
    e0.record(tr.stream)
    for i in range(a.iters):
        tr.step(i % a.ring)
    e1.record(tr.stream)
    tr.synchronize()
    res["step_us"] = round(1e3 * e0.elapsed_time(e1) / a.iters, 2)
    # one graph per pass over the ring (programmatic dependent launch between the steps)
    ring = list(range(a.ring))
    tr.capture_epoch(ring)
    for _ in range(3):
        tr.run_epoch(ring)
    reps = max(1, a.iters // a.ring)
    e0.record(tr.stream)
    for _ in range(reps):
        tr.run_epoch(ring)
    e1.record(tr.stream)
    tr.synchronize()
    res["epoch_step_us"] = round(1e3 * e0.elapsed_time(e1) / (reps * a.ring), 2)
    res["gnodes_per_s"] = round(res["nodes"] / min(res["step_us"], res["epoch_step_us"]) / 1e3, 3)

    print(json.dumps(res), flush=True)


if __name__ == "__main__":
    main()
