#!/usr/bin/env python
"""Block-level view of an `ncu --page source --csv --print-source sass` export: consecutive SASS
instructions with the same execution count are one block; prints the blocks that matter with their
opcode mix and stall samples.   usage: sass_blocks.py sass.csv [kernel-index] [min-share]"""
import collections
import csv
import sys

rows = list(csv.reader(open(sys.argv[1])))
want = int(sys.argv[2]) if len(sys.argv) > 2 else 0
min_share = float(sys.argv[3]) if len(sys.argv) > 3 else 0.004
starts = [k for k, r in enumerate(rows) if r and r[0] == "Kernel Name"]
k0 = starts[want]
k1 = starts[want + 1] if want + 1 < len(starts) else len(rows)
print(rows[k0][1][:100])
hdr = rows[k0 + 1]
ci = {h: k for k, h in enumerate(hdr)}
stall_cols = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
grp = []
for r in rows[k0 + 2:k1]:
    try:
        ie = int(r[ci["Instructions Executed"]])
        sm = int(r[ci["# Samples"]])
    except (ValueError, IndexError):
        continue
    toks = r[ci["Source"]].split()
    op = toks[1] if toks and toks[0].startswith("@") and len(toks) > 1 else (toks[0] if toks else "")
    op = op.split(".")[0]
    st = collections.Counter({h[6:]: int(r[ci[h]] or 0) for h in stall_cols})
    if grp and grp[-1][0] == ie:
        grp[-1][1] += 1
        grp[-1][2] += sm
        grp[-1][3][op] += 1
        grp[-1][4].update(st)
    else:
        grp.append([ie, 1, sm, collections.Counter({op: 1}), st])
tot = sum(g[0] * g[1] for g in grp)
tots = sum(g[2] for g in grp)
print("total warp instructions", tot, "samples", tots, "static instructions", sum(g[1] for g in grp))
for g in grp:
    if g[0] * g[1] > min_share * tot:
        stalls = {k: v for k, v in g[4].most_common(5) if v}
        print(f"exec {g[0]:>8} x {g[1]:>4} instr = {100 * g[0] * g[1] / tot:5.1f}% inst, {100 * g[2] / max(tots, 1):5.1f}% samples | "
              f"{dict(g[3].most_common(10))} | {stalls}")
