"""Pins the CPU oracle against the REFERENCE'S OWN hot path executed here.

`oracle/_ref/libnekcem_ref.so` (recipe: oracle/build_ref.py) contains the reference's Fortran
routines from cem_maxwell_op_rk downwards, translated statement by statement by
oracle/f2c_lite.py from the sources under /root/reference/src, linked with the reference's own
gather-scatter library src/jl/gs.c compiled unchanged.  Both sides get the same COMMON-block
inputs and run the reference's shipped test cases; the oracle's hand-written restatement must
reproduce the translated reference BIT FOR BIT (both are compiled without FMA contraction), for
the fields, the RK registers and the PML / ADE auxiliary state.

On a machine without /root/reference the prebuilt library is used if it travelled with the
tree; otherwise these tests are skipped.
"""
import ctypes as C

import numpy as np
import pytest

from oracle import cases, refrun

pytestmark = pytest.mark.skipif(not refrun.available(), reason="oracle/_ref not built and no reference tree")


def _pair(case):
    return case, refrun.ReferenceRun(case)


def _assert_same(case, ref, what="fields"):
    n3 = 3 * case.npts
    for name in ("hn", "en", "khn", "ken"):
        a, b = getattr(case, name), ref.view(name)[:n3]
        assert np.array_equal(a, b), (what, name, float(np.abs(a - b).max()))
    if case.ifpml:
        for name in ("pmlbn", "pmldn", "kpmlbn", "kpmldn"):
            a, b = getattr(case, name), ref.view(name)[:n3]
            assert np.array_equal(a, b), (what, name, float(np.abs(a - b).max()))
    assert np.abs(case.hn).max() + np.abs(case.en).max() > 1e-8  # not a comparison of zeros


def test_units_translated_from_reference():
    names = refrun.lib().ref_units().decode().split()
    for must in ("cem_maxwell_op_rk", "maxwell_wght_curl", "local_grad3", "mxm", "mxf9",
                 "cem_maxwell_flux3d", "cem_maxwell_flux2d", "cem_maxwell_add_flux_to_res",
                 "pml_step", "rk4_upd", "rk_storage", "cem_maxwell_drude", "cem_maxwell_lorentz"):
        assert must in names


def test_rk_coefficients_are_the_references():
    """rk_storage (src/cem_common.F:78-114): the oracle's LSRK(5,4) table equals the one the
    translated reference computes from its rational literals."""
    c, r = _pair(cases.case_boxper((2, 2, 2), 4))
    for k in ("rk4a", "rk4b", "rk4c"):
        assert np.array_equal(np.array(getattr(c.s, k)), r.view(k)[:len(getattr(c.s, k))]), k
    r.close()


def test_pin_3dboxper_as_shipped():
    """tests/3dboxper (128 elements from the reference's .re2, N=8, dt=2e-4), 20 steps."""
    c, r = _pair(cases.case_3dboxper())
    for _ in range(4):
        c.step(5); r.step(5)
        _assert_same(c, r)
    assert r.get("time") == c.time
    r.close()


def test_pin_stage_by_stage():
    """every RK stage separately, incl. rktime and the residual arrays after cem_maxwell_op"""
    c, r = _pair(cases.case_boxper((3, 2, 2), 6))
    for st in range(1, 6):
        c.stage(st); r.stage(st)
        assert r.get("rktime") == c.s.rktime
        for name in ("reshn", "resen"):
            assert np.array_equal(getattr(c, name), r.view(name)[:3 * c.npts]), (st, name)
        _assert_same(c, r, "stage %d" % st)
    r.close()


@pytest.mark.parametrize("nx1", [2, 3, 5, 8, 12, 16, 17])
def test_pin_orders(nx1):
    """mxm dispatches to a different unrolled mxfK for every order (src/nek5_mxm_std.F)"""
    c, r = _pair(cases.case_boxper((2, 2, 2), nx1))
    c.step(2); r.step(2)
    _assert_same(c, r)
    r.close()


def test_pin_3dboxpec():
    """tests/3dboxpec: PEC walls -> cem_maxwell_flux_pec and the doubled impedances"""
    c, r = _pair(cases.case_3dboxpec())
    c.step(20); r.step(20)
    _assert_same(c, r)
    r.close()


def test_pin_central_flux():
    from oracle import oracle as O
    mesh = O.box_mesh((2, 2, 2), ((0.0, 2 * np.pi),) * 3, ("P  ",) * 6)
    c = O.RefCase(mesh, 5, upwind=False)
    c.set_dt(-1e-3)
    shn, sen = cases.usersol_3dboxper(c, 0.0)
    c.hn[:] = shn; c.en[:] = sen
    r = refrun.ReferenceRun(c)
    c.step(3); r.step(3)
    _assert_same(c, r)
    r.close()


@pytest.mark.parametrize("twomat", [False, True])
def test_pin_3ddielectric(twomat):
    """tests/3ddielectric: PML (pml_step + the PML half of rk_maxwell_ab), heterogeneous
    eps/mu, userinc plane-wave injection between restrict_to_face and the flux"""
    c = cases.case_3ddielectric(twomat)
    r = refrun.ReferenceRun(c)
    r.set_callback("userinc", c.user.userinc(c))
    c.step(10); r.step(10)
    _assert_same(c, r)
    r.close()


def test_pin_3dboxpml():
    """tests/3dboxpml: all-PML box, usersrc dipole between pml_step and invqmass"""
    c = cases.case_3dboxpml(nx1=6, nel=(5, 5, 5))
    r = refrun.ReferenceRun(c)
    r.set_callback("usersrc", cases.usersrc_3dboxpml(c))
    c.step(15); r.step(15)
    _assert_same(c, r)
    r.close()


@pytest.mark.parametrize("imode", [1, 2])
def test_pin_2dboxper(imode):
    """tests/2dboxper TE / TM: local_grad2, flux2d"""
    c, r = _pair(cases.case_2dboxper(imode))
    c.step(50); r.step(50)
    _assert_same(c, r)
    r.close()


@pytest.mark.parametrize("imode", [1, 2])
def test_pin_2dboxpec(imode):
    c, r = _pair(cases.case_2dboxpec(imode))
    c.step(50); r.step(50)
    _assert_same(c, r)
    r.close()


@pytest.mark.parametrize("kind", ["drude", "lorentz"])
def test_pin_dispersive(kind):
    """tests/drude, tests/lorentz: 2D TE + PEC + PML + userinc + the ADE advanced by the
    reference's cem_maxwell_drude / cem_maxwell_lorentz called from usersrc"""
    c = cases.case_drude() if kind == "drude" else cases.case_lorentz()
    u = c.user
    r = refrun.ReferenceRun(c)
    r.set_callback("userinc", u.userinc(c))
    jn, kjn, resjn = u.jn.copy(), u.kjn.copy(), u.resjn.copy()
    params = u.params.copy()
    index1 = (u.index + 1).astype(np.int32)
    n = C.c_int(index1.size)
    fn = r.L.cem_maxwell_drude_ if kind == "drude" else r.L.cem_maxwell_lorentz_
    dp = lambda a: a.ctypes.data_as(C.POINTER(C.c_double))

    def usersrc(tt, *res):  # drude.usr:153-170
        fn(dp(jn), dp(kjn), dp(resjn), dp(params), index1.ctypes.data_as(C.POINTER(C.c_int)),
           C.byref(n))

    r.set_callback("usersrc", usersrc)
    c.step(40); r.step(40)
    _assert_same(c, r)
    assert np.array_equal(u.jn, jn) and np.array_equal(u.kjn, kjn)
    assert np.abs(jn).max() > 1e-3
    r.close()


def test_pin_cemface_numbering():
    """cem_set_fc_ptr (src/cem_common.F:214-283) translated from the reference fills cemface
    from skpdat; the oracle's face->volume map must be identical."""
    from oracle import oracle as O
    for ldim, nx1, nel in ((3, 4, (2, 1, 2)), (2, 5, (2, 3))):
        mesh = O.box_mesh(nel, ((0.0, 1.0),) * ldim, ("P  ",) * (2 * ldim))
        c = O.RefCase(mesh, nx1, imode=1)
        r = refrun.ReferenceRun(c)
        # eface / skpdat exactly as setup_topo fills them: initds, dsset(nx1,ny1,nz1)
        # (src/nek5_connect11.F:1046-1093, 1440-1528), both translated from the reference
        nz1 = nx1 if ldim == 3 else 1
        r.L.initds_()
        r.L.dsset_(C.byref(C.c_int(nx1)), C.byref(C.c_int(nx1)), C.byref(C.c_int(nz1)))
        r.view("cemface")[:] = 0
        r.L.cem_set_fc_ptr_()
        assert r.get("ncemface") == c.nxzfl
        assert np.array_equal(r.view("cemface")[:c.nxzfl], c.cemface + 1)
        r.close()
