import sys, numpy as np
sys.path.insert(0, "/root/repo"); sys.path.insert(0, "/root/repo/tests")
from helpers import rel_l2
import test_gpu_zgraphene as T
from oracle import cases
for inc in (True, False):
    for strict in (False, True):
        c = cases.case_2dgraphene(1)
        if not inc:
            c.s.userinc = type(c.s.userinc)()
        s = T._solver(c, incident=inc, strict=strict)
        s.step(200); c.step(200)
        (fg, fo), (kg, ko) = T._sheet_state(c, s)
        print("incident", inc, "strict", strict, "fields %.2e fj %.2e kj %.2e" % (rel_l2(T._fields(s), T._fields(c)), rel_l2(fg, fo), rel_l2(kg, ko)),
              "bitwise fields", np.array_equal(T._fields(s), T._fields(c)), "max|f|", np.abs(T._fields(c)).max(), flush=True)
        s.close()
