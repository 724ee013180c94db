"""The remaining Maxwell configurations of the reference's test suite on the GPU:
tests/2ddielectric (TE/TM x one/two materials: PML + incident hook in the 2D kernel) and
tests/2dboxpml (TE/TM: all-PML box + the volume-source hook on hz / ez), tests/cylwave (curved
unstructured mesh)."""
import numpy as np
import pytest

from helpers import rel_l2, solver_from_refcase

pytestmark = pytest.mark.gpu
TOL = 1e-12


def _fields(obj):
    return np.concatenate([obj.hn, obj.en])


@pytest.mark.parametrize("imode", [1, 2])
@pytest.mark.parametrize("twomat", [False, True])
def test_kat_2ddielectric_on_gpu(imode, twomat):
    """as shipped: parity with the oracle (fields, PML fields) after 200 steps and the .usr
    tolerances at steps 1..10, 100, 200 through the device-side cem_error"""
    from oracle import cases
    c = cases.case_2ddielectric(imode, twomat)
    s = solver_from_refcase(c, incident=c.user.incident(c))
    done = 0
    for target in list(range(1, 11)) + [100, 200]:
        s.step(target - done); c.step(target - done)
        done = target
        shn, sen = c.usersol(c, s.time)
        l2, linf = s.cem_error(shn, sen)
        assert np.all(l2 <= np.array(c.tol["l2"]) + 1e-300), (target, l2)
        assert np.all(linf <= np.array(c.tol["linf"]) + 1e-300), (target, linf)
    assert rel_l2(_fields(s), _fields(c)) <= TOL
    assert rel_l2(s.get_array("pmlbn"), c.pmlbn) <= TOL
    assert rel_l2(s.get_array("pmldn"), c.pmldn) <= TOL
    s.close()


@pytest.mark.parametrize("imode", [1, 2])
def test_2dboxpml_with_gaussian_source(imode):
    """as shipped (8x8 elements, all PML, CFL 0.1): the .usr's usersrc registered as a separable
    volume source on hz (TE) / ez (TM); parity after 60 steps and the userchk bound"""
    from oracle import cases
    c = cases.case_2dboxpml(imode)
    s = solver_from_refcase(c)
    cb = c.usersrc_fn
    # usersrc: res(comp) -= profile * (sin(-omega t) * bm1)
    s.set_volume_source(cb.comp, cb.profile, amp=1.0, omega=-cb.omega, phase=0.0)
    s.step(60); c.step(60)
    assert np.abs(_fields(c)).max() > 1e-4
    assert rel_l2(_fields(s), _fields(c)) <= TOL
    assert rel_l2(s.get_array("pmlbn"), c.pmlbn) <= TOL
    assert rel_l2(s.get_array("pmldn"), c.pmldn) <= TOL
    assert np.abs(_fields(s)).max() < 1.0
    s.close()


def test_kat_cylwave_on_gpu():
    """tests/cylwave as shipped (the reference's unstructured 50-element mesh with circular-arc
    sides, N=11, PEC wall, periodic in z): truly curved elements through the general-metric path
    (the setup scan must find no constant-metric element), parity with the oracle after 100 steps
    and the .usr tolerances 5e-9 / 5e-8 at steps 1..10 and 100."""
    from oracle import cases
    c = cases.case_cylwave()
    s = solver_from_refcase(c)
    ncm, _ = s.geometry_info()
    assert ncm == 0
    done = 0
    for target in list(range(1, 11)) + [100]:
        s.step(target - done); c.step(target - done)
        done = target
        shn, sen = c.usersol(c, s.time)
        l2, linf = s.cem_error(shn, sen)
        assert np.all(l2 <= 5e-9) and np.all(linf <= 5e-8), (target, l2, linf)
    assert rel_l2(_fields(s), _fields(c)) <= TOL
    s.close()
