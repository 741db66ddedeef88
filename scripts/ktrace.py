#!/usr/bin/env python
"""Phase timeline of the one-launch train kernel (Args::trace marks): per-phase durations, median
over CTAs, for one workload.   python scripts/ktrace.py --mesh 30 30 --batch 256"""
import argparse
import ctypes as C
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import torch


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--mesh", type=int, nargs="+", default=[30, 30])
    ap.add_argument("--batch", type=int, default=256)
    ap.add_argument("--layers", type=int, default=4)
    a = ap.parse_args()
    from g_adaptivity_b200 import GNN, _lib, synth
    from g_adaptivity_b200.trainer import DeformerTrainer
    md = tuple(a.mesh)
    opt = synth.default_opt(md)
    opt.update(num_layers=a.layers, device="cuda:0", gad_store_alpha=False)
    ds = synth.SyntheticDataset(len(md), md)
    torch.manual_seed(42)
    model = GNN(ds, opt).to("cuda:0")
    tr = DeformerTrainer(model, use_cuda_graph=False)
    sids = [tr.add_batch(synth.make_batch(md, a.batch, seed=100 + r)) for r in range(3)]
    lib = _lib.load()
    trace = torch.zeros((1024, 64), dtype=torch.int64, device="cuda:0")
    with torch.cuda.stream(tr.stream):
        for rep in range(4):
            for sid in sids:
                s = tr.slots[sid]
                d = tr._train_desc(s, 2)
                if rep == 3 and sid == sids[-1]:
                    trace.zero_()
                    d.trace = trace.data_ptr()
                _lib.check(lib.gad_train_step_ell(C.byref(d), tr.stream.cuda_stream), "train")
    tr.synchronize()
    t = trace.cpu().numpy()
    T = tr.slots[0].graph.T
    grid = int((t[:, 0] > 0).sum())
    t = t[:grid]
    L = a.layers
    red = None
    if grid == T + 1:            # reducer CTA (owns no tile): entry, dry pass, all tiles in, reduced, chain rule, Adam, done
        red, t = t[grid - 1], t[:grid - 1]
    names = ["entry->inputs", "assemble->pdl_wait"] + [f"fwd{l}" for l in range(L)] + ["loss reduce", "backward+reduce"]
    t0 = t[:, 0].min()
    nm = 2 + L + 2 + 1
    d = np.diff(t[:, :nm], axis=1) / 1e3
    print(f"grid {grid} CTAs, {T} tiles; tile CTAs span {(t[:, :nm].max() - t0) / 1e3:.2f} us")
    print("entry spread (us): min %.2f med %.2f max %.2f" % tuple(np.percentile((t[:, 0] - t0) / 1e3, [0, 50, 100])))
    for k, n in enumerate(names):
        print(f"  {n:20s} med {np.median(d[:, k]):7.2f}  p10 {np.percentile(d[:, k], 10):7.2f}  p90 {np.percentile(d[:, k], 90):7.2f} us")
    end = t[:, nm - 1]
    print("CTA end (us after first entry): med %.2f max %.2f" % (np.median(end - t0) / 1e3, (end.max() - t0) / 1e3))
    if red is not None:
        r = (red[:7] - t0) / 1e3
        print("reducer CTA: entry %.2f | prepared %.2f | all tiles in %.2f (last tile CTA end %.2f) | reduced +%.2f | "
              "chain rule +%.2f | adam +%.2f | refold +%.2f -> done at %.2f us" %
              (r[0], r[1], r[2], (end.max() - t0) / 1e3, r[3] - r[2], r[4] - r[3], r[5] - r[4], r[6] - r[5], r[6]))
    else:
        k = int(np.argmax(t[:, nm + 4]))
        if t[k, nm + 4] > 0:
            tl = t[k, nm - 1:nm + 5]
            print("tail CTA %d: own end %.2f | elected +%.2f | reduced +%.2f | chain rule +%.2f | adam +%.2f | refold +%.2f -> done at %.2f us" %
                  ((k, (tl[0] - t0) / 1e3) + tuple(np.diff(tl) / 1e3) + ((tl[-1] - t0) / 1e3,)))


if __name__ == "__main__":
    main()
