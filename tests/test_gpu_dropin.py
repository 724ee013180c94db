"""The drop-in boundary exercised the way the reference would use it.

`oracle/_ref/libnekcem_ref_dropin.so` holds the reference's own routines (translated from
/root/reference/src by oracle/f2c_lite.py) together with THIS REPO'S Fortran shim
`fortran/cem_maxwell_b200_f77.F`, translated by the same tool and linked against
`libnekcem_b200.so`.  The only interface between the two sides is the reference's COMMON blocks
(SIZE/TOTAL/EMWAVE/PML, /bdry1/ cempec, /c_is1/ glo_num): the shim's `b200_copy_all_in`
(replacing acc_copy_all_in, src/cem_drive.F:397-480) uploads them through the `_` twins of the C
ABI, `b200_op_rk` replaces `call cem_maxwell_op_rk` (src/cem_drive.F:628), `b200_update_host`
is the `!$ACC UPDATE HOST` seam.  The same COMMON state is then advanced by the reference's own
cem_maxwell_op_rk, and the fields must agree within 1e-12 relative L2.
"""
import ctypes as C

import numpy as np
import pytest

from helpers import rel_l2

pytestmark = pytest.mark.gpu
TOL = 1e-12


def _refrun():
    from oracle import refrun
    if not refrun.available("dropin"):
        pytest.skip("oracle/_ref/libnekcem_ref_dropin.so did not travel with the tree")
    return refrun


def _time_loop(r, call, nsteps):
    """src/cem_drive.F:618-654: istep loop, `call <op_rk>`, time = time + dt"""
    for _ in range(nsteps):
        call()
        r.set("time", r.get("time") + r.get("dt"))


def _run_both(case, nsteps, callbacks=None, names=("hn", "en")):
    refrun = _refrun()
    n3 = 3 * case.npts
    # (1) the reference's own path
    r = refrun.ReferenceRun(case, kind="dropin")
    for which, fn in (callbacks or {}).items():
        r.set_callback(which, fn("ref"))
    _time_loop(r, r.L.cem_maxwell_op_rk_, nsteps)
    want = {k: r.view(k)[:n3].copy() for k in names}
    t_end = r.get("time")
    r.close()
    # (2) the same COMMON blocks driven through the shim into the GPU library
    r = refrun.ReferenceRun(case, kind="dropin")
    for which, fn in (callbacks or {}).items():
        r.set_callback(which, fn("gpu"))
    r.L.b200_copy_all_in_()
    r.L.b200_update_device_()
    for k in ("hn", "en"):
        r.view(k)[:] = -7.0                      # the host copies must come back from the device
    _time_loop(r, r.L.b200_op_rk_, nsteps)
    r.L.b200_update_host_()
    got = {k: r.view(k)[:n3].copy() for k in ("hn", "en")}
    assert r.get("time") == t_end
    r.L.b200_copy_all_out_()
    r.close()
    a = np.concatenate([got["hn"], got["en"]]); b = np.concatenate([want["hn"], want["en"]])
    assert np.abs(b).max() > 1e-3
    return rel_l2(a, b)


def test_dropin_3dboxper_as_shipped():
    """tests/3dboxper (mesh from the reference's .re2): 20 steps through the Fortran shim"""
    from oracle import cases
    assert _run_both(cases.case_3dboxper(), 20) <= TOL


def test_dropin_3dboxpec():
    """PEC walls: cempec (1-based face points from COMMON /bdry1/) and the doubled impedances"""
    from oracle import cases
    assert _run_both(cases.case_3dboxpec(), 20) <= TOL


@pytest.mark.parametrize("imode", [1, 2])
def test_dropin_2dboxper(imode):
    from oracle import cases
    assert _run_both(cases.case_2dboxper(imode), 40) <= TOL


def test_dropin_pml_two_materials():
    """tests/3ddielectric geometry, materials and PML (pmlptr, pmlsigma, pmlbn/pmldn from COMMON
    /pml1-3/); the userinc injection is left out on both sides (the shim has no device-side
    registration for it; tests/test_gpu_parity.py covers the incident hook)"""
    from oracle import cases
    assert _run_both(cases.case_3ddielectric(True), 15) <= TOL


def test_dropin_drude_through_the_users_usersrc():
    """tests/drude: the .usr's usersrc calls cem_maxwell_drude(jn,kjn,resjn,params,dindex,n)
    (drude.usr:153-170).  In the drop-in build that name resolves to the library's twin: the
    first call (made by b200_update_device) registers the user's arrays, the ADE then advances
    inside the fused kernel.  Reference side: the oracle, which equals the translated reference
    bit for bit on this case (tests/test_reference_pin.py::test_pin_dispersive).  userinc is
    left out on both sides (no device-side registration in the shim)."""
    from oracle import cases
    refrun = _refrun()
    nsteps = 20
    c = cases.case_drude()
    c.set_callback("userinc", lambda tt, *a: None)
    c.step(nsteps)

    c2 = cases.case_drude()
    u = c2.user
    drop = refrun.lib("dropin")      # its cem_maxwell_drude_ is the product library's twin
    idx1 = (u.index + 1).astype(np.int32)
    n = C.c_int(idx1.size)
    dp = lambda a: a.ctypes.data_as(C.POINTER(C.c_double))
    jn, kjn, resjn, par = u.jn.copy(), u.kjn.copy(), u.resjn.copy(), u.params.copy()

    def usersrc(tt, *res):
        drop.cem_maxwell_drude_(dp(jn), dp(kjn), dp(resjn), dp(par),
                                idx1.ctypes.data_as(C.POINTER(C.c_int)), C.byref(n))

    r = refrun.ReferenceRun(c2, kind="dropin")
    r.set_callback("usersrc", usersrc)
    r.L.b200_copy_all_in_()
    r.L.b200_update_device_()
    _time_loop(r, r.L.b200_op_rk_, nsteps)
    r.L.b200_update_host_()
    n3 = 3 * c.npts
    got = np.concatenate([r.view("hn")[:n3], r.view("en")[:n3]])
    r.L.b200_copy_all_out_()
    r.close()
    assert rel_l2(got, np.concatenate([c.hn, c.en])) <= TOL
    assert np.abs(c.user.jn).max() > 1e-3


def test_dropin_incident_field_registered_like_a_usr_would():
    """tests/3ddielectric complete: the reference side runs the .usr's userinc callback between
    restrict_to_face and the flux (src/cem_maxwell.F:498); the drop-in side registers the same
    plane wave once through the Fortran twin nekcem_b200_set_incident (as a .usr would do in
    usrdat2) and the injection happens inside the fused kernel."""
    from helpers import incident_3ddielectric
    from oracle import cases
    refrun = _refrun()
    nsteps = 15
    c = cases.case_3ddielectric(True)
    n3 = 3 * c.npts
    r = refrun.ReferenceRun(c, kind="dropin")
    r.set_callback("userinc", c.user.userinc(c))
    _time_loop(r, r.L.cem_maxwell_op_rk_, nsteps)
    want = np.concatenate([r.view("hn")[:n3], r.view("en")[:n3]]).copy()
    r.close()

    r = refrun.ReferenceRun(c, kind="dropin")
    r.L.b200_copy_all_in_()
    j, amp, phase, omega = incident_3ddielectric(c)
    fp1 = np.ascontiguousarray(j + 1, dtype=np.int32)          # 1-based, like incindex
    amp = np.ascontiguousarray(amp, dtype=np.float64); phase = np.ascontiguousarray(phase)
    h = C.c_int(int(r.get("b200_handle"))); ninc = C.c_int(fp1.size); om = C.c_double(omega)
    dp = lambda a: a.ctypes.data_as(C.POINTER(C.c_double))
    r.L.nekcem_b200_set_incident_(C.byref(h), C.byref(ninc), fp1.ctypes.data_as(C.POINTER(C.c_int)),
                                  dp(amp), dp(phase), C.byref(om))
    r.L.b200_update_device_()
    _time_loop(r, r.L.b200_op_rk_, nsteps)
    r.L.b200_update_host_()
    got = np.concatenate([r.view("hn")[:n3], r.view("en")[:n3]]).copy()
    r.L.b200_copy_all_out_()
    r.close()
    assert rel_l2(got, want) <= TOL


def test_dropin_volume_source_registered_like_a_usr_would():
    """tests/3dboxpml: the reference side runs the .usr's usersrc (Gaussian dipole,
    3dboxpml.usr:30-88) between pml_step and invqmass; the drop-in side registers profile,
    amplitude and frequency once through nekcem_b200_set_volume_source_."""
    from oracle import cases
    refrun = _refrun()
    nsteps = 15
    c = cases.case_3dboxpml(nx1=6, nel=(5, 5, 5))
    n3 = 3 * c.npts
    fn = cases.usersrc_3dboxpml(c)
    r = refrun.ReferenceRun(c, kind="dropin")
    r.set_callback("usersrc", fn)
    _time_loop(r, r.L.cem_maxwell_op_rk_, nsteps)
    want = np.concatenate([r.view("hn")[:n3], r.view("en")[:n3]]).copy()
    r.close()

    r = refrun.ReferenceRun(c, kind="dropin")
    r.L.b200_copy_all_in_()
    h = C.c_int(int(r.get("b200_handle")))
    prof = np.ascontiguousarray(fn.profile, dtype=np.float64)
    comp, amp, om, ph = C.c_int(5), C.c_double(1.0), C.c_double(-fn.omega), C.c_double(0.0)
    r.L.nekcem_b200_set_volume_source_(C.byref(h), C.byref(comp),
                                       prof.ctypes.data_as(C.POINTER(C.c_double)),
                                       C.byref(amp), C.byref(om), C.byref(ph))
    r.L.b200_update_device_()
    _time_loop(r, r.L.b200_op_rk_, nsteps)
    r.L.b200_update_host_()
    got = np.concatenate([r.view("hn")[:n3], r.view("en")[:n3]]).copy()
    r.L.b200_copy_all_out_()
    r.close()
    assert np.abs(want).max() > 1e-6
    assert rel_l2(got, want) <= TOL


def test_dropin_restart_handoff_through_the_shim():
    """The reference's `maxwell-restart` test (tests/restart/restart.usr:70-165) through the
    Fortran shim and the COMMON blocks: HN, EN = 1.0 go to the device, b200_restart_out delivers the
    payloads cem_out / cem_restart_out would write, the fields are overwritten with 2.0 on host and
    device, b200_restart_swap (the field part of restart_swap) reads the payloads back, and after
    b200_update_host the COMMON arrays must hold 1.0 again -- the .usr demands cem_error <= 1e-15;
    here every bit."""
    from oracle import cases
    refrun = _refrun()
    case = cases.case_boxper((3, 3, 3), 8, dt=-1e-3)
    n3 = 3 * case.npts
    r = refrun.ReferenceRun(case, kind="dropin")
    if not hasattr(r.L, "b200_restart_swap_"):
        pytest.skip("oracle/_ref was built before the shim had the restart routines")
    for k in ("hn", "en"):
        r.view(k)[:n3] = 1.0
    r.L.b200_copy_all_in_()
    r.L.b200_update_device_()
    buf_e = np.zeros(n3); buf_h = np.zeros(n3)
    dbl = C.c_int(1)
    dp = C.POINTER(C.c_double)
    r.L.b200_restart_out_(buf_e.ctypes.data_as(dp), buf_h.ctypes.data_as(dp), C.byref(dbl))
    assert buf_e.tobytes() == np.ones(n3).astype(">f8").tobytes()      # big-endian 1.0
    for k in ("hn", "en"):
        r.view(k)[:n3] = 2.0
    r.L.b200_update_device_()                                            # the device holds 2.0
    r.L.b200_restart_swap_(buf_e.ctypes.data_as(dp), buf_h.ctypes.data_as(dp), C.byref(dbl))
    for k in ("hn", "en"):
        r.view(k)[:n3] = -7.0
    r.L.b200_update_host_()
    for k in ("hn", "en"):
        assert np.array_equal(r.view(k)[:n3], np.ones(n3)), k
    r.L.b200_copy_all_out_()
    r.close()
