"""Developer sweep: one periodic box per (order, elems), timed for several option settings.
Usage: python scripts/sweep.py "N:E[:steps]" ... [--opt pf_dist=0,148,740]"""
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np  # noqa: E402
from nekcem_b200 import MaxwellB200  # noqa: E402
from nekcem_b200.boxcase import BoxCase  # noqa: E402

PEAK = 6459.0
try:
    PEAK = float(json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"])
except Exception:
    pass

cfgs, opts = [], {"pf_dist": [0]}
for a in sys.argv[1:]:
    if a.startswith("--opt"):
        continue
    if "=" in a:
        k, v = a.split("=")
        opts[k] = [int(x) for x in v.split(",")]
    else:
        cfgs.append([int(x) for x in a.split(":")])
for cfg in cfgs:
    order, E = cfg[0], cfg[1]
    steps = cfg[2] if len(cfg) > 2 else 5
    nx1 = order + 1
    case = BoxCase((E, E, E), nx1)
    s = MaxwellB200(3, nx1, case.nelt, device=0)
    s.cem_maxwell_init(case.lazy(), free_after_upload=True)
    s.setup()
    s.set_time(0.0, 1e-4)
    bytes_stage = s.algorithmic_bytes_per_stage()
    for k, vals in opts.items():
        for v in vals:
            s.set_option(k, v)
            s.step(2)
            best = 1e30
            for _ in range(3):
                s.step(steps)
                ms, nl = s.last_step_ms()
                best = min(best, ms / (5 * steps))
            rate = case.npts / (best * 1e-3) / 1e9
            print(f"N={order} E={E} {k}={v}: {best:.3f} ms/stage  {rate:.2f} Gnode-stage/s  "
                  f"roofline {bytes_stage / (best * 1e-3) / 1e9 / PEAK:.3f}", flush=True)
    shn, sen = case.fields(s.time)
    l2, linf = s.cem_error(shn, sen)
    print(f"   l2 err vs analytic {l2.max():.2e}", flush=True)
    s.close()
