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


@pytest.mark.parametrize("which", ["3ddielectric-2mat", "2ddielectric-tm", "drude", "lorentz",
                                   "3dgraphene", "2dgraphene-te"])
def test_device_side_usersol_planewave(which):
    """usersol of the layered-media tests evaluated on the device
    (nekcem_b200_error_sums_planewave: two half-space plane waves with complex amplitudes and
    wavenumbers + the graded PML decay).  The L2 / Linf errors must equal those of cem_error
    against the host usersol (to the round-off of device vs host exp/sin/cos/pow) and meet the
    .usr tolerances; no exact-solution array crosses PCIe."""
    from helpers import planewave_args
    from oracle import cases
    import test_gpu_zgraphene as G
    mk = {"3ddielectric-2mat": lambda: cases.case_3ddielectric(True),
          "2ddielectric-tm": lambda: cases.case_2ddielectric(2, True),
          "drude": cases.case_drude, "lorentz": cases.case_lorentz,
          "3dgraphene": lambda: cases.case_3dgraphene(nel=(3, 12, 3)),
          "2dgraphene-te": lambda: cases.case_2dgraphene(1)}[which]
    c = mk()
    u = c.user
    if "graphene" in which:
        s = G._solver(c)
    elif which in ("drude", "lorentz"):
        s = solver_from_refcase(c, incident=u.incident(c),
                                ade=(which, u.jn, u.kjn, u.params, u.index))
    elif which.startswith("3d"):
        from helpers import incident_3ddielectric
        s = solver_from_refcase(c, incident=incident_3ddielectric(c))
    else:
        s = solver_from_refcase(c, incident=u.incident(c))
    s.set_array("ymn", c.ym1)
    s.step(10)
    tt = s.time
    a = planewave_args(c)
    l2d, linfd = s.cem_error_planewave(a["omega"], a["k"], a["amp"], a["region"], a["inpml"],
                                       a["pml"], tt)
    shn, sen = c.usersol(c, tt)
    l2h, linfh = s.cem_error(shn, sen)
    assert np.abs(l2h).max() > 1e-12                      # a real truncation error is visible
    assert np.allclose(l2d, l2h, rtol=1e-6, atol=1e-15), (l2d, l2h)
    assert np.allclose(linfd, linfh, rtol=1e-6, atol=1e-14), (linfd, linfh)
    assert np.all(l2d <= np.array(c.tol["l2"]) + 1e-300)
    assert np.all(linfd <= np.array(c.tol["linf"]) + 1e-300)
    s.close()


@pytest.mark.parametrize("case", ["3d", "2d-te"])
def test_vtk_payload_assembled_on_device(case):
    """Output hand-off: the VTK "VECTORS" payload of EN and HN -- per node
    three values, cast to float32 (or kept as float64), big-endian -- assembled on the device must
    be byte-identical to the oracle's restatement of vtk_nonswap_field + writefield4[_double],
    which is pinned to the reference's own pieces (tests/test_reference_pin.py)."""
    from oracle import cases
    c = cases.case_boxper((3, 3, 3), 5) if case == "3d" else cases.case_2dboxper(1, nx1=6)
    s = solver_from_refcase(c)
    s.step(3)
    # the oracle restates the byte format; feed it the GPU's own fields so that the comparison
    # is about the payload, bit for bit, not about the time stepping
    c.hn[:] = s.hn; c.en[:] = s.en
    for which in ("en", "hn"):
        for dbl in (False, True):
            got = s.vtk_payload(which, as_double=dbl)
            want = c.vtk_payload(which, as_double=dbl)
            assert len(got) == 3 * c.npts * (8 if dbl else 4)
            assert got == want, (which, dbl)
    assert np.abs(c.en).max() > 1e-3
    s.close()
