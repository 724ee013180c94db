"""cem_maxwell_op as a stand-alone operator (nekcem_b200_apply_rhs): what the reference's
exponential / eigenvalue drivers call through amult (src/cem_maxwell.F:2310-2365).  Against the
oracle's cem_maxwell_op (reshn, resen after invqmass).

Host-side plumbing around the fused stage: one launch with (a, b, dt) = (0, 0, 1).  Green on B200
since the round-1 driver run."""
import ctypes as C

import numpy as np
import pytest

from helpers import rel_l2, solver_from_refcase

pytestmark = [pytest.mark.gpu]
TOL = 1e-12


@pytest.mark.parametrize("which", ["3dboxper", "3dboxpec", "2dboxper-te", "3ddielectric"])
def test_rhs_operator(which):
    from oracle import cases
    c = {"3dboxper": cases.case_3dboxper, "3dboxpec": cases.case_3dboxpec,
         "2dboxper-te": lambda: cases.case_2dboxper(1),
         "3ddielectric": lambda: cases.case_3ddielectric(True)}[which]()
    if which == "3ddielectric":
        c.set_callback("userinc", lambda tt, *a: None)
    s = solver_from_refcase(c)
    s.step(3); c.step(3)                        # a state with non-trivial RK registers
    h0, e0 = s.hn.copy(), s.en.copy()
    t = c.s.time + 0.3 * c.s.dt
    c.s.rkstep = 1
    c.s.rktime = t
    c.L.ora_cem_maxwell_op(C.byref(c.s))
    rh, re = s.cem_maxwell_op(t)
    assert np.abs(c.reshn).max() > 1e-3
    assert rel_l2(np.concatenate([rh, re]), np.concatenate([c.reshn, c.resen])) <= TOL
    assert np.array_equal(s.hn, h0) and np.array_equal(s.en, e0)     # fields untouched
    s.close()
