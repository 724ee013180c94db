"""Pins the CPU oracle with the reference's own known-answer tests: the analytic solutions
and the L2/Linf tolerances compiled into tests/<case>/<case>.usr.  The
meshes of 3dboxper / 3dboxpec come from the reference's .re2/.rea/.map files via the
committed fixtures in tests/golden/ (generator: tests/golden/make_golden.py)."""
import numpy as np
import pytest

from oracle import cases


def _check(c, steps_to_check, nsteps):
    done = 0
    for target in steps_to_check:
        c.step(target - done)
        done = target
        l2, linf = c.errors(c.usersol)
        for k in range(6):
            assert l2[k] <= c.tol["l2"][k], (target, k, l2[k])
            assert linf[k] <= c.tol["linf"][k], (target, k, linf[k])
    assert done == nsteps
    return l2, linf


def test_kat_3dboxper_full():
    """tests/3dboxper: 128 elements, N=8, 50 steps of dt=2e-4; userchk at steps 1..10 and 50
    with L2 <= 5e-10, Linf <= 5e-9 (3dboxper.usr:181-212)."""
    c = cases.case_3dboxper()
    assert c.nelt == 128 and c.nx1 == 9 and c.npts == 93312
    l2, linf = _check(c, list(range(1, 11)) + [50], 50)
    assert max(l2) > 1e-12  # the scheme's truncation error is visible: not a trivial pass


def test_kat_3dboxpec_full():
    """tests/3dboxpec: 27 elements, PEC walls, 1000 steps of dt=5e-3; L2 <= 5e-8, Linf <= 5e-7
    (3dboxpec.usr:198-228), checked every iocomm=10... here every 100 steps."""
    c = cases.case_3dboxpec()
    assert c.ifpec and not c.ifpml and c.ncempec == 9 * 6 * 81
    _check(c, list(range(100, 1001, 100)), 1000)


@pytest.mark.parametrize("twomat", [False, True])
def test_kat_3ddielectric(twomat):
    """tests/3ddielectric (1 and 2 materials): PML in +-y, plane-wave injection through
    userinc; tolerances 5e-4 / 5e-3 on hx,hz,ex,ez and ~1e-14 / 1e-12 on the zero components
    hy, ey (3ddielectric.usr userchk).  300 of the 1000 steps (the error is periodic)."""
    c = cases.case_3ddielectric(twomat)
    assert c.ifpml and c.maxpml == 64 and c.user.incindex.size == 16 * 81
    for target in (100, 200, 300):
        c.step(100)
        l2, linf = c.errors(c.usersol)
        for k in (0, 2, 3, 5):
            assert l2[k] <= 5e-4 and linf[k] <= 5e-3, (target, k, l2[k], linf[k])
        for k in (1, 4):
            assert l2[k] <= 5e-14 and linf[k] <= 5e-12, (target, k, l2[k], linf[k])


def test_kat_3dboxpml_stability():
    """tests/3dboxpml: all-PML box with a Gaussian dipole; userchk only requires the fields to
    stay below 1.0.  200 of the 2000 steps at CFL 0.1."""
    c = cases.case_3dboxpml(nx1=7, nel=(6, 6, 6))
    assert c.maxpml == 6 ** 3 - 4 ** 3
    c.step(200)
    assert np.all(np.isfinite(c.hn)) and np.all(np.isfinite(c.en))
    assert np.max(np.abs(c.en)) < 1.0 and np.max(np.abs(c.hn)) < 1.0
    assert np.max(np.abs(c.en)) > 1e-6  # the source did radiate


@pytest.mark.parametrize("imode", [1, 2])
def test_kat_2dboxper(imode):
    """tests/2dboxper TE (param(4)=1) and TM (=2): 9 elements, N=8, 1000 steps of dt=5e-3;
    userchk at steps 1..10 and every 100 with L2 <= 5e-8, Linf <= 5e-7 on the three active
    components and exact zeros elsewhere (2dboxper.usr:199-231)."""
    c = cases.case_2dboxper(imode)
    assert c.ldim == 2 and c.nelt == 9 and c.npts == 9 * 81
    _check(c, list(range(1, 11)) + list(range(100, 1001, 100)), 1000)


@pytest.mark.parametrize("imode", [1, 2])
def test_kat_2dboxpec(imode):
    """tests/2dboxpec TE/TM: PEC walls, tolerances 1e-6 / 1e-5 (2dboxpec.usr:204-231)."""
    c = cases.case_2dboxpec(imode)
    assert c.ifpec and c.ncempec == 12 * 9
    _check(c, list(range(1, 11)) + list(range(100, 1001, 100)), 1000)


def test_kat_drude():
    """tests/drude: 2D TE, 4x32 elements, PEC bottom / PML top, plane-wave injection, Drude ADE
    through usersrc -> cem_maxwell_drude; 1e-7 / 5e-6 on hz, ex, ey at steps 1..10 and every 50
    (drude.usr userchk).  400 of the 1000 steps."""
    c = cases.case_drude()
    assert c.nelt == 128 and c.maxpml == 24 and c.user.index.size == 64 * 81
    _check(c, list(range(1, 11)) + list(range(50, 401, 50)), 400)
    assert np.max(np.abs(c.user.jn)) > 0.1  # the current is alive


def test_kat_lorentz():
    """tests/lorentz: as drude with the two-pole ADE; 5e-6 / 5e-5 on hz, ex and 5e-10 on ey
    (lorentz.usr:373-387).  400 of the 1000 steps."""
    c = cases.case_lorentz()
    _check(c, list(range(1, 11)) + list(range(100, 401, 100)), 400)
    n = c.npts
    assert np.max(np.abs(c.user.jn[3 * n:4 * n])) > 1e-3  # the polarisation variable too


@pytest.mark.parametrize("imode", [1, 2])
def test_kat_2dgraphene(imode):
    """tests/2dgraphene TE (param(4)=1) and TM (=2): 4x32 elements, N=8, CFL 0.2, PML in +-y,
    plane wave onto a graphene sheet at y=0 whose surface current (Drude + two critical-point
    terms) is advanced per face point by userfsrc -> cem_te/tm_graphene_current; 1e-7 / 5e-6 on
    the two wave components and 1e-14 / 5e-13 on the zero one, userchk at steps 1..10 and every
    100 (2dgraphene.usr userchk).  All 1000 steps."""
    c = cases.case_2dgraphene(imode)
    assert c.nelt == 128 and c.user.graphindex.size == 8 * 9 and c.user.incindex.size == 4 * 9
    _check(c, list(range(1, 11)) + list(range(100, 1001, 100)), 1000)
    assert np.max(np.abs(c.user.fjn)) > 1e-2  # the sheet carries current


def test_kat_3dgraphene():
    """tests/3dgraphene: 4x12x4 elements, N=8, dt=5e-3, TE and TM waves superimposed; 5e-4 / 5e-3
    on hx,hz,ex,ez and 1e-14 / 5e-12 on hy,ey (3dgraphene.usr:483-495), iocomm = 50.  300 of the
    1000 steps (the error is periodic; the full run was checked once: max L2 1.8e-4, hy 6.5e-15)."""
    c = cases.case_3dgraphene()
    assert c.nelt == 192 and c.user.graphindex.size == 32 * 81 and c.user.incindex.size == 16 * 81
    _check(c, list(range(1, 11)) + list(range(50, 301, 50)), 300)


@pytest.mark.parametrize("imode", [1, 2])
@pytest.mark.parametrize("twomat", [False, True])
def test_kat_2ddielectric(imode, twomat):
    """tests/2ddielectric TE/TM with one and two materials: 4x32 elements, N=8, dt=5e-3, PML
    thick 10 in +-y, plane-wave injection through userinc; tolerances of the .usr's userchk
    (5e-8 / 1e-6 one material, 5e-7 / 5e-6 two; ~1e-14 on the zero component) at steps 1..10 and
    every 100 of all 1000 steps."""
    c = cases.case_2ddielectric(imode, twomat)
    assert c.nelt == 128 and c.maxpml == 80 and c.user.incindex.size == 4 * 9
    _check(c, list(range(1, 11)) + list(range(100, 1001, 100)), 1000)


@pytest.mark.parametrize("imode", [1, 2])
def test_kat_2dboxpml_stability(imode):
    """tests/2dboxpml TE/TM: all-PML 8x8 box (thick 2) with a Gaussian source; userchk only
    requires the fields to stay below 1.0.  800 of the 4000 steps at CFL 0.1."""
    c = cases.case_2dboxpml(imode)
    assert c.maxpml == 64 - 16
    c.step(800)
    assert np.all(np.isfinite(c.hn)) and np.all(np.isfinite(c.en))
    assert np.max(np.abs(c.en)) < 1.0 and np.max(np.abs(c.hn)) < 1.0
    assert np.max(np.abs(c.en)) > 1e-3  # the source did radiate


def test_kat_cylwave():
    """tests/cylwave: TM_01 mode of a circular PEC waveguide on the reference's own unstructured
    mesh (50 hexahedra from cylwave.rea/.map, 80 circular-arc sides generated by the restated
    ARCSRF), N=11, CFL 0.25, periodic in z; 5e-9 / 5e-8 on all six components at steps 1..10 and
    every 100 (cylwave.usr userchk).  300 of the 1000 steps."""
    c = cases.case_cylwave()
    assert c.nelt == 50 and c.nx1 == 12 and c.ifpec and c.ncempec == 40 * 144
    # curved elements: the quadrature reproduces the cylinder's volume pi r^2 L
    assert abs(c.volvm1 - np.pi * c.cyl_radius ** 2 * c.zm1.max()) < 1e-11 * c.volvm1
    _check(c, list(range(1, 11)) + [100, 200, 300], 300)
