"""Drop-in checks with a padded SIZE layout (lelt = nelt + 3), run in a process of their own by
tests/test_gpu_zz_padded_size.py: the Fortran twins exit(1) on any library error -- the reference's
error behaviour -- which must not take the test runner down with it.
Usage: python tests/drivers/padded_dropin_check.py pml|drude      (exit code 0 = parity within 1e-12)"""
import ctypes as C
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
from helpers import rel_l2  # noqa: E402

TOL = 1e-12


def _refrun():
    from oracle import refrun
    if not refrun.available("dropin"):
        print("SKIP: oracle/_ref/libnekcem_ref_dropin.so did not travel with the tree")
        sys.exit(77)
    return refrun


def check_pml():
    """3ddielectric geometry with PML (pmlsigma, pmlbn, pmldn are (lpts1,3) too)"""
    from oracle import cases
    refrun = _refrun()
    c = cases.case_3ddielectric(True, nx1=6, nel=(3, 6, 3))
    c.set_callback("userinc", lambda tt, *a: None)
    r = refrun.ReferenceRun(c, kind="dropin", pad_elems=3)
    c.step(10)
    r.L.b200_copy_all_in_()
    r.L.b200_update_device_()
    for k in ("hn", "en"):
        r.view(k)[:] = -7.0
    for _ in range(10):
        r.L.b200_op_rk_()
        r.set("time", r.get("time") + r.get("dt"))
    r.L.b200_update_host_()
    got = np.concatenate([r.field("hn"), r.field("en")])
    # the padding of the host arrays is left alone
    assert np.all(r.view("hn")[c.npts:r.lpts1] == -7.0)
    r.L.b200_copy_all_out_()
    r.close()
    assert rel_l2(got, np.concatenate([c.hn, c.en])) <= TOL


def check_drude():
    """tests/drude with jn(lpts,3), params(lpts,2) dimensioned by the padded SIZE, registered
    through the .usr's usersrc -> cem_maxwell_drude twin"""
    from oracle import cases
    refrun = _refrun()
    nsteps = 10
    c = cases.case_drude()
    c.set_callback("userinc", lambda tt, *a: None)
    c.step(nsteps)
    c2 = cases.case_drude()
    u = c2.user
    r = refrun.ReferenceRun(c2, kind="dropin", pad_elems=3)
    lp, n = r.lpts1, c2.npts
    drop = refrun.lib("dropin")
    pad = lambda a, m: np.concatenate([np.concatenate([a[k * n:(k + 1) * n], np.zeros(lp - n)])
                                       for k in range(m)])
    jn, kjn, resjn, par = pad(u.jn, 3), pad(u.kjn, 3), pad(u.resjn, 3), pad(u.params, 2)
    idx1 = (u.index + 1).astype(np.int32)
    nn = C.c_int(idx1.size)
    dp = lambda a: a.ctypes.data_as(C.POINTER(C.c_double))

    def usersrc(tt, *res):
        drop.cem_maxwell_drude_(dp(jn), dp(kjn), dp(resjn), dp(par),
                                idx1.ctypes.data_as(C.POINTER(C.c_int)), C.byref(nn))

    r.set_callback("usersrc", usersrc)
    r.L.b200_copy_all_in_()
    r.L.b200_update_device_()
    for _ in range(nsteps):
        r.L.b200_op_rk_()
        r.set("time", r.get("time") + r.get("dt"))
    r.L.b200_update_host_()
    got = np.concatenate([r.field("hn"), r.field("en")])
    h = C.c_int(int(r.get("b200_handle")))
    drop.nekcem_b200_get_ade_(C.byref(h), dp(jn), dp(kjn))
    r.L.b200_copy_all_out_()
    r.close()
    assert rel_l2(got, np.concatenate([c.hn, c.en])) <= TOL
    jc = np.concatenate([jn[k * lp:k * lp + n] for k in range(3)])
    assert np.abs(c.user.jn).max() > 1e-3 and rel_l2(jc, c.user.jn) <= TOL


if __name__ == "__main__":
    {"pml": check_pml, "drude": check_drude}[sys.argv[1]]()
    print("ok", sys.argv[1])
