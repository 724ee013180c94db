"""Developer sweep: one periodic box per (order, elems), timed for every combination of the
option values given.
Usage: python scripts/sweep.py "N:E[:steps]" ... [const_metrics=0,1] [pipeline=0,1] [pre:xtrace=0]
       [case:pml=all] [case:eps_upper=4.0] [case:dim=2]   (BoxCase variants / the 2D TE box)"""
import itertools
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np  # noqa: E402,F401
from nekcem_b200 import MaxwellB200  # noqa: E402
from nekcem_b200.boxcase import BoxCase, BoxCase2D  # noqa: E402

PEAK = 6459.0
try:
    PEAK = float(json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"])
except Exception:
    pass

cfgs, opts, pre, variant = [], {}, {}, {}
for a in sys.argv[1:]:
    if a.startswith("case:"):
        k, v = a[5:].split("=")
        variant[k] = float(v) if k == "eps_upper" else (int(v) if k == "dim" else v)
    elif a.startswith("pre:"):  # applied before setup (e.g. pre:xtrace=0)
        k, v = a[4:].split("=")
        pre[k] = int(v)
    elif "=" in a:
        k, v = a.split("=")
        opts[k] = [int(x) for x in v.split(",")]
    else:
        cfgs.append([int(x) for x in a.split(":")])
if not opts:
    opts = {"const_metrics": [0]}
keys = list(opts)
for cfg in cfgs:
    order, E = cfg[0], cfg[1]
    steps = cfg[2] if len(cfg) > 2 else 5
    nx1 = order + 1
    dim = variant.pop("dim", 3) if "dim" in variant else 3
    if dim == 2:
        case = BoxCase2D((E, E), nx1, imode=1)
        s = MaxwellB200(2, nx1, case.nelt, imode=1, device=0)
        s.cem_maxwell_init(case.arrays())
    else:
        case = BoxCase((E, E, E), nx1, **variant)
        s = MaxwellB200(3, nx1, case.nelt, device=0, ifpml=bool(variant.get("pml")))
        s.cem_maxwell_init(case.lazy(), free_after_upload=True)
    for k, v in pre.items():
        s.set_option(k, v)
    s.setup()
    s.set_time(0.0, 1e-4)
    bytes_stage = s.algorithmic_bytes_per_stage()
    for combo in itertools.product(*[opts[k] for k in keys]):
        for k, v in zip(keys, combo):
            s.set_option(k, v)
        s.step(2)
        best = 1e30
        for _ in range(3):
            s.step(steps)
            ms, nl = s.last_step_ms()
            best = min(best, ms / (5 * steps))
        rate = case.npts / (best * 1e-3) / 1e9
        tag = " ".join([f"{k}={v}" for k, v in zip(keys, combo)] + [f"{k}={v}" for k, v in pre.items()])
        print(f"N={order} E={E} {tag}: {best:.3f} ms/stage  {rate:.2f} Gnode-stage/s  "
              f"roofline(B(n)) {bytes_stage / (best * 1e-3) / 1e9 / PEAK:.3f}", flush=True)
    if not variant:
        shn, sen = case.fields(s.time)
        l2, linf = s.cem_error(shn, sen)
        print(f"   l2 err vs analytic {l2.max():.2e}", flush=True)
    s.close()
