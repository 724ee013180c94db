import sys, numpy as np
sys.path.insert(0, "/root/repo"); sys.path.insert(0, "/root/repo/tests")
from helpers import rel_l2, solver_from_refcase
from oracle import cases
for nx1 in (11, 12, 13, 16):
    for pipeline in (0, 1):
        c = cases.case_boxper((3, 3, 4), nx1, dt=-1e-3)
        s = solver_from_refcase(c)
        s.set_option("pipeline", pipeline)
        c.step(3); s.step(3)
        f = rel_l2(np.concatenate([s.get_array("hn"), s.get_array("en")]), np.concatenate([c.hn, c.en]))
        kg = np.concatenate([s.get_array("khn"), s.get_array("ken")]); ko = np.concatenate([c.khn, c.ken])
        per = [rel_l2(kg[i*c.npts:(i+1)*c.npts], ko[i*c.npts:(i+1)*c.npts]) for i in range(6)]
        print(nx1, pipeline, "fields", f, "k", rel_l2(kg, ko), "percomp", ["%.1e" % p for p in per], "absmax k", np.abs(ko).max(), flush=True)
        s.close()
