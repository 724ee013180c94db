"""The optional modal filter at the end of every time step (`if (iffilter) call
q_filter(0.01)`, src/cem_maxwell.F:342, src/nek5_filter.F): filter_kernel against the oracle,
whose filterq is pinned bit for bit to the reference's (tests/test_reference_pin.py)."""
import ctypes as C

import numpy as np
import pytest

from helpers import rel_l2, solver_from_refcase

pytestmark = pytest.mark.gpu
TOL = 1e-12


def _fields(obj):
    return np.concatenate([obj.hn, obj.en])


@pytest.mark.parametrize("which", ["3d-n7", "3d-n16", "3d-n24", "2d-te", "2d-tm"])
def test_filtered_time_stepping(which):
    from oracle import cases
    if which == "3d-n7":
        mk = lambda: cases.case_boxper((3, 3, 3), 7, dt=-2e-3)
    elif which == "3d-n16":
        mk = lambda: cases.case_boxper((3, 3, 3), 16, dt=-5e-4)
    elif which == "3d-n24":  # the largest element: 226 KB of shared memory in filter_kernel
        mk = lambda: cases.case_boxper((3, 3, 3), 24, dt=-2e-4)
    else:
        mk = lambda: cases.case_2dboxper(1 if which.endswith("te") else 2, nx1=8)
    c = mk()
    c.set_filter()
    s = solver_from_refcase(c)
    s.set_filter(c.filter)
    s.step(6); c.step(6)
    assert rel_l2(_fields(s), _fields(c)) <= TOL
    ms, launches = s.last_step_ms()
    assert launches >= 36 and launches % 6 == 0    # per step: five stages per element list + one filter launch
    if which not in ("3d-n16", "3d-n24"):
        # the filter is not a no-op here (at N=15 this smooth mode has nothing in the top two
        # Legendre modes: that case exercises the 67 KB shared-memory configuration only)
        c0 = mk()
        c0.step(6)
        assert rel_l2(_fields(c0), _fields(c)) > 1e-9
    # switching it off restores the unfiltered scheme
    s.set_filter(None)
    s.step(1); c.filter = None; c.step(1)
    assert rel_l2(_fields(s), _fields(c)) <= TOL
    s.close()


def test_dropin_filter_through_the_shim():
    """iffilter = .true. in COMMON /INPUT/: the shim's b200_copy_all_in builds the matrix with the
    reference's own build_new_filter and registers it (nekcem_b200_set_filter_); b200_op_rk then
    filters on the device.  Reference side: the oracle with the same (translated) matrix -- its
    filterq equals the reference's bit for bit."""
    from oracle import cases, refrun
    if not refrun.available("dropin"):
        pytest.skip("oracle/_ref/libnekcem_ref_dropin.so did not travel with the tree")
    nsteps = 8
    c = cases.case_3dboxper()
    r = refrun.ReferenceRun(c, kind="dropin")
    n = c.nx1
    intv = np.zeros(n * n)
    z = np.ascontiguousarray(c.zgm1)
    dp = lambda a: a.ctypes.data_as(C.POINTER(C.c_double))
    r.L.build_new_filter_(dp(intv), dp(z), C.byref(C.c_int(n)), C.byref(C.c_int(2)),
                          C.byref(C.c_double(0.01)), C.byref(C.c_int(1)))
    c.filter = intv
    c.step(nsteps)
    r.set("iffilter", 1)
    r.put("zgm1", np.concatenate([c.zgm1, c.zgm1, c.zgm1]))
    r.set("nid_io", 1)                             # keep build_new_filter quiet
    r.L.b200_copy_all_in_()
    r.L.b200_update_device_()
    for _ in range(nsteps):
        r.L.b200_op_rk_()
        r.set("time", r.get("time") + r.get("dt"))
    r.L.b200_update_host_()
    n3 = 3 * c.npts
    got = np.concatenate([r.view("hn")[:n3], r.view("en")[:n3]])
    r.L.b200_copy_all_out_()
    r.close()
    assert rel_l2(got, np.concatenate([c.hn, c.en])) <= TOL
