"""cfg 4 (one 200x200 mesh, 64 RK4 steps, forward) on both streaming routes: the cooperative persistent kernel and the
chain of dependent launches (GAD_WIDE_PERSIST=0).  python scripts/cfg4_bench.py [n ...]"""
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402

if __name__ == "__main__":
    dev = torch.device("cuda:0")
    peak = 6539.9
    sizes = [int(a) for a in sys.argv[1:]] or [200]
    for n in sizes:
        for persist in ("1", "0"):
            os.environ["GAD_WIDE_PERSIST"] = persist
            r = bench.time_forward_config(dev, (n, n), 1, {"ode_method": "rk4", "num_layers": 64}, False, 20, peak)
            print(json.dumps({"n": n, "persist": persist, "module_ms": r["module_call"]["ms"],
                              "session_ms": r["session_replay"]["ms"], "frac": r["session_replay"]["roofline_frac"]}), flush=True)
