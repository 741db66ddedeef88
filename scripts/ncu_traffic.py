#!/usr/bin/env python
"""Per-launch DRAM traffic of each kernel in an `ncu --set full` report -> profiles/traffic.json
(read by bench.py for `roofline.traffic`).   usage: ncu_traffic.py report.ncu-rep [out.json]"""
import csv
import io
import json
import os
import subprocess
import sys

rep = sys.argv[1]
out = sys.argv[2] if len(sys.argv) > 2 else os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))),
                                                         "profiles", "traffic.json")
txt = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(txt)))
hdr, units = rows[0], rows[1]
ci = {h: k for k, h in enumerate(hdr)}
scale = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
acc = {}
for r in rows[2:]:
    name = r[ci["Kernel Name"]].split("<")[0].split("(")[0].split("::")[-1].replace("void ", "").strip()
    tot = 0.0
    for col in ("dram__bytes_read.sum", "dram__bytes_write.sum"):
        tot += float(r[ci[col]]) * scale[units[ci[col]]]
    dur_us = float(r[ci["gpu__time_duration.sum"]]) * {"us": 1.0, "ns": 1e-3, "ms": 1e3}[units[ci["gpu__time_duration.sum"]]]
    a = acc.setdefault(name, {"launches": 0, "dram": 0.0, "us": 0.0})
    a["launches"] += 1
    a["dram"] += tot
    a["us"] += dur_us
res = {k: {"dram_bytes_per_launch": v["dram"] / v["launches"], "ncu_duration_us": v["us"] / v["launches"],
           "launches_captured": v["launches"], "report": os.path.basename(rep)} for k, v in acc.items()}
if os.path.exists(out):
    old = json.load(open(out))
    old.update(res)
    res = old
json.dump(res, open(out, "w"), indent=1, sort_keys=True)
print(json.dumps(res, indent=1))
