"""Developer timing of the auxiliary 3D paths at benchmark size (all-PML box, two materials + PML layers)."""
import sys, os, json
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT)
import numpy as np
from nekcem_b200 import MaxwellB200
from nekcem_b200.boxcase import BoxCase
PEAK = float(json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"])
for order, E, var in [(7, 40, dict(pml="all")), (7, 40, dict(pml="layers", eps_upper=4.0)), (8, 32, dict(pml="all"))]:
    nx1 = order + 1
    case = BoxCase((E, E, E), nx1, **var)
    s = MaxwellB200(3, nx1, case.nelt, device=0, ifpml=True)
    s.cem_maxwell_init(case.lazy(), free_after_upload=True)
    s.set_option("const_metrics", 0)
    s.setup(); s.set_time(0.0, 1e-4)
    npml = case.array("pmlptr").size
    bpn = 280 + 696.0 / nx1 + 240.0 * npml / case.nelt
    s.step(2); best = 1e30
    for _ in range(3):
        s.step(3); ms, _ = s.last_step_ms(); best = min(best, ms / 15)
    print(f"N={order} E={E} {var}: {best:.3f} ms/stage {case.npts/best/1e6:.2f} Gnode-stage/s frac {bpn*case.npts/(best*1e-3)/1e9/PEAK:.3f} ({bpn:.0f} B/node)", flush=True)
    s.close()
