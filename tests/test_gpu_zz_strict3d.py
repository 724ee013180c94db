"""desc.strict = 1 in 3D contexts: the slab stage kernel compiled a second time with -fmad=false
(nekcem_b200/csrc/Makefile: _obj/stage_slab_strict.o), so that every product and sum is rounded
separately as in the reference's x86-64 build (DESIGN.md "Numerics").  What the mode shows: the
distance between the default build and the oracle is FMA contraction plus the association of the
six-term curl sum -- both <= 1e-15 per operation -- and the two builds of the library agree with
each other and with the oracle far inside the 1e-12 bar, at low, middle and the highest order."""
import numpy as np
import pytest

from helpers import incident_3ddielectric, rel_l2, solver_from_refcase

pytestmark = pytest.mark.gpu
TOL = 1e-12


def _state(s):
    return np.concatenate([s.get_array(k) for k in ("hn", "en", "khn", "ken")])


def _fields(c):
    return np.concatenate([c.hn, c.en])


@pytest.mark.parametrize("nx1", [5, 8, 9, 12, 16, 20])
def test_strict_3d_periodic_box_vs_oracle(nx1):
    from oracle import cases
    c = cases.case_boxper((3, 3, 3), nx1, dt=-1e-3)
    out = {}
    for strict in (False, True):
        s = solver_from_refcase(c, strict=strict)
        s.step(3)
        out[strict] = _state(s)
        s.close()
    c.step(3)
    ref = np.concatenate([c.hn, c.en, c.khn, c.ken])
    nf = 2 * c.hn.size
    for strict in (False, True):
        assert rel_l2(out[strict][:nf], ref[:nf]) <= TOL, (strict, nx1)
        # the RK registers are dt*res, a difference of O(N^2) larger terms (test_gpu_pipe.py)
        assert rel_l2(out[strict][nf:], ref[nf:]) <= (TOL if nx1 <= 10 else 1e-10), (strict, nx1)
    # the two builds differ by rounding only
    assert rel_l2(out[True][:nf], out[False][:nf]) <= 1e-13
    print(f"nx1={nx1}: fields vs oracle default {rel_l2(out[False][:nf], ref[:nf]):.2e} "
          f"strict {rel_l2(out[True][:nf], ref[:nf]):.2e}; RK registers default "
          f"{rel_l2(out[False][nf:], ref[nf:]):.2e} strict {rel_l2(out[True][nf:], ref[nf:]):.2e}")


def test_strict_3d_dielectric_pml_incident():
    """tests/3ddielectric (two materials, PML elements: the auxiliary instantiation, incident
    plane wave) through the strict build"""
    from oracle import cases
    c = cases.case_3ddielectric(twomat=True)
    s = solver_from_refcase(c, incident=incident_3ddielectric(c), strict=True)
    c.step(5); s.step(5)
    assert rel_l2(np.concatenate([s.hn, s.en]), _fields(c)) <= TOL
    s.close()
