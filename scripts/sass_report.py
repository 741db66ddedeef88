#!/usr/bin/env python
"""Digest of `ncu -i X.ncu-rep --page source --csv --print-source sass`: stall mix, opcode mix and the
most-sampled instructions of each profiled kernel.  Usage: sass_report.py sass.csv [kernel-substring]"""
import collections
import csv
import sys


def num(x):
    try:
        return int(float(x))
    except ValueError:
        return 0


def main():
    rows = list(csv.reader(open(sys.argv[1])))
    want = sys.argv[2] if len(sys.argv) > 2 else ""
    kernels, cur = [], None
    for r in rows:
        if r and r[0] == "Kernel Name":
            cur = {"name": r[1], "hdr": None, "data": []}
            kernels.append(cur)
        elif cur is not None and cur["hdr"] is None and r and r[0] == "Address":
            cur["hdr"] = r
        elif cur is not None and cur["hdr"] is not None and len(r) == len(cur["hdr"]):
            cur["data"].append(r)
    seen = set()
    for k in kernels:
        short = k["name"].split("(")[0]
        if want not in short or short in seen:
            continue
        seen.add(short)
        idx = {h: i for i, h in enumerate(k["hdr"])}
        data = k["data"]
        ti = sum(num(r[idx["Instructions Executed"]]) for r in data)
        ts = sum(num(r[idx["# Samples"]]) for r in data)
        print(f"=== {short}: {len(data)} SASS lines, {ti} warp-instructions, {ts} samples")
        stalls = [h for h in k["hdr"] if h.startswith("stall_") and "Not Issued" not in h]
        agg = {s: sum(num(r[idx[s]]) for r in data) for s in stalls}
        print("  stalls: " + ", ".join(f"{s[6:]} {100 * v / max(ts, 1):.1f}%" for s, v in sorted(agg.items(), key=lambda kv: -kv[1])[:8]))
        mix = collections.Counter()
        for r in data:
            toks = [t for t in r[idx["Source"]].split() if not t.startswith("@")]
            if toks:
                mix[toks[0].split(".")[0]] += num(r[idx["Instructions Executed"]])
        print("  opcodes: " + ", ".join(f"{o} {100 * v / max(ti, 1):.1f}%" for o, v in mix.most_common(18)))
        exc = sum(num(r[idx["L1 Wavefronts Shared Excessive"]]) for r in data)
        wav = sum(num(r[idx["L1 Wavefronts Shared"]]) for r in data)
        print(f"  shared wavefronts {wav}, excessive {exc} ({100 * exc / max(wav, 1):.1f}%)")
        print("  most-sampled instructions:")
        for r in sorted(data, key=lambda r: -num(r[idx["# Samples"]]))[:18]:
            print(f"    {num(r[idx['# Samples']]):6d} smp {num(r[idx['Instructions Executed']]):8d} exe  {r[idx['Source']][:100]}")


if __name__ == "__main__":
    main()
