"""C-ABI library: loads without a GPU, exports every symbol include/nekcem_b200.h declares,
host-side planning works in a host-only context, compute entry points fail loudly."""
import ctypes as C
import os
import re

import numpy as np
import pytest

from nekcem_b200 import MaxwellB200, NekcemB200Error, lib
from nekcem_b200.api import ARRAY_IDS, LIBPATH
from nekcem_b200.boxcase import BoxCase

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_library_exports_every_declared_symbol():
    hdr = open(os.path.join(ROOT, "include", "nekcem_b200.h")).read()
    names = set(re.findall(r"\b(nekcem_b200_[a-z0-9_]+)\s*\(", hdr))
    assert len(names) >= 25
    L = C.CDLL(LIBPATH)
    for n in sorted(names):
        assert hasattr(L, n), f"{n} declared in the header but not exported"
    for n in ("nekcem_b200_create_", "nekcem_b200_step_", "nekcem_b200_set_array_",
              "nekcem_b200_get_array_", "nekcem_b200_set_faces_", "nekcem_b200_set_pml_",
              "nekcem_b200_setup_", "nekcem_b200_set_time_", "nekcem_b200_set_incident_",
              "nekcem_b200_set_volume_source_", "nekcem_b200_error_sums_",
              "nekcem_b200_error_sums_mode_", "nekcem_b200_error_sums_planewave_",
              "nekcem_b200_comm_unique_id_", "nekcem_b200_comm_init_",
              "nekcem_b200_set_drude_", "nekcem_b200_set_lorentz_", "nekcem_b200_get_ade_",
              "nekcem_b200_bind_", "cem_maxwell_drude_", "cem_maxwell_lorentz_",
              "nekcem_b200_set_graphene_", "nekcem_b200_get_graphene_",
              "nekcem_b200_set_filter_", "nekcem_b200_set_rk_coefficients_",
              "nekcem_b200_vtk_payload_", "nekcem_b200_sync_host_", "nekcem_b200_sync_device_",
              "cem_3d_graphene_current_", "cem_te_graphene_current_",
              "cem_tm_graphene_current_", "nekcem_b200_step_streamed_",
              "nekcem_b200_restart_ingest_", "nekcem_b200_device_count_",
              "nekcem_b200_set_option_"):
        assert hasattr(L, n), f"Fortran twin {n} missing"


def test_array_enum_matches_header():
    hdr = open(os.path.join(ROOT, "include", "nekcem_b200.h")).read()
    body = hdr[hdr.index("enum nekcem_b200_array"):]
    body = body[body.index("{") + 1:body.index("NKB_ARRAY_COUNT")]
    body = re.sub(r"/\*.*?\*/", "", body, flags=re.S)
    ids = re.findall(r"NKB_([A-Z0-9_]+)", body)
    assert [i.lower() for i in ids] == [k.lower() for k in ARRAY_IDS]


def test_bad_arguments_fail_loudly():
    with pytest.raises(NekcemB200Error, match="nx1"):
        MaxwellB200(3, 40, 8, device=-1)
    # the reference's own range of orders: mxf1..mxf24 (src/nek5_mxm_std.F)
    with pytest.raises(NekcemB200Error, match="nx1"):
        MaxwellB200(3, 25, 8, device=-1)
    with pytest.raises(NekcemB200Error, match="nx1"):
        MaxwellB200(2, 1, 8, imode=1, device=-1)
    for ldim, imode in ((3, 3), (2, 1)):
        MaxwellB200(ldim, 24, 2, imode=imode, device=-1, strict=True).close()
    with pytest.raises(NekcemB200Error, match="imode"):
        MaxwellB200(2, 8, 8, imode=3, device=-1)
    with pytest.raises(NekcemB200Error, match="imode"):
        MaxwellB200(3, 8, 8, imode=1, device=-1)
    s = MaxwellB200(3, 4, 27, device=-1)
    # a host-only context keeps what is uploaded (so that the upload path can be inspected) and
    # refuses every compute call
    v = np.arange(s.npts, dtype=np.float64)
    s.set_array("rxmn", v)
    assert np.array_equal(s.get_array("rxmn"), v)
    with pytest.raises(NekcemB200Error, match="never set"):
        s.get_array("rymn")
    with pytest.raises(NekcemB200Error, match="host-only"):
        s.set_volume_source(5, np.zeros(s.npts), 1.0, 1.0, 0.0)
    with pytest.raises(NekcemB200Error, match="set_faces has not been called"):
        s.setup()
    with pytest.raises(NekcemB200Error, match="nxzfl"):
        s.set_faces(np.zeros(5, dtype=np.int64), np.zeros(0))
    s.close()
    L = lib()
    assert L.nekcem_b200_step(12345, 1) != 0
    assert b"invalid" in L.nekcem_b200_last_error()


def test_no_cpu_fallback_without_device():
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    with pytest.raises(NekcemB200Error, match="no CUDA device|CPU fallback"):
        MaxwellB200(3, 4, 27, device=0)


def test_host_plan_periodic_box():
    b = BoxCase((3, 4, 5), 4)
    s = MaxwellB200(3, 4, b.nelt, device=-1)
    s.set_faces(b.array("glo_num"), b.array("cempec"))
    vm, peers, nhalo, ni, nb = s.plan()
    assert not peers and nhalo == 0 and ni == b.nelt and nb == 0
    assert vm.min() >= 0  # every face point has a local partner
    # the partner's partner is the own node: build own node index per face point
    n, nfp = 4, 6 * 16
    own = np.zeros(s.nxzfl, dtype=np.int64)
    for e in range(b.nelt):
        for f in range(6):
            for p in range(16):
                a, bb = p % n, p // n
                node = [a + 16 * bb, 3 + 4 * a + 16 * bb, a + 12 + 16 * bb, 4 * a + 16 * bb,
                        a + 4 * bb, a + 4 * bb + 48][f]
                own[e * nfp + f * 16 + p] = e * 64 + node
    inv = {int(o): j for j, o in enumerate(own)}  # face nodes are shared by up to 3 faces ...
    x, y, z = b.coords()
    L = b.length
    for arr in (x, y, z):
        d = np.abs(arr[own] - arr[vm])
        d = np.minimum(d, np.abs(d - L))
        assert d.max() < 1e-12
    s.close()


def test_host_plan_pec_box_codes():
    b = BoxCase((2, 2, 2), 3, bc="PEC")
    s = MaxwellB200(3, 3, b.nelt, device=-1)
    g = b.array("glo_num")
    s.set_faces(g, b.array("cempec"))
    vm = s.plan()[0]
    assert np.all(vm[g == 0] == -1) and np.all(vm[g != 0] >= 0)
    s.set_faces(g, np.zeros(0))  # unpaired and not PEC -> code -2
    vm = s.plan()[0]
    assert np.all(vm[g == 0] == -2)
    s.close()


def test_host_plan_2d_box():
    """2D face numbering: slot s, point p of an element sits at node (p,0), (n-1,p), (p,n-1),
    (0,p) -- the reference's cemface (src/cem_common.F:262-283) as restated by the oracle."""
    from oracle import cases
    c = cases.case_2dboxper(1, nx1=5, nel=(3, 4))
    s = MaxwellB200(2, 5, c.nelt, imode=1, device=-1)
    s.set_faces(c.glo_num, np.zeros(0))
    vm = s.plan()[0]
    assert vm.min() >= 0
    # partner of partner: vmapP of the partner face point is the own node = cemface
    own = c.cemface.astype(np.int64)
    inv = {}
    for j, o in enumerate(own):
        inv.setdefault(int(o), []).append(j)
    for j in range(vm.size):
        partners = inv[int(vm[j])]
        assert any(vm[q] == own[j] for q in partners)
    # geometric check: partner nodes coincide modulo the period
    for arr in (c.xm1, c.ym1):
        d = np.abs(arr[own] - arr[vm])
        d = np.minimum(d, np.abs(d - 2 * np.pi))
        assert d.max() < 1e-12
    s.close()


def test_graphene_registration_host_side():
    """set_graphene / get_graphene in a host-only context: the (nxzfl,3,6) user arrays are packed
    into the library's per-sheet-point layout and come back at the listed face points only;
    bad indices fail loudly."""
    from oracle import cases
    c = cases.case_2dgraphene(1, nx1=4, nel=(3, 6))
    u = c.user
    s = MaxwellB200(2, 4, c.nelt, imode=1, device=-1)
    s.set_faces(c.glo_num, c.cempec[:c.ncempec])
    rng = np.random.default_rng(3)
    fjn = rng.standard_normal(18 * c.nxzfl); kfjn = rng.standard_normal(18 * c.nxzfl)
    s.cem_graphene_current(fjn, kfjn, u.graphparams, c.yconduc, u.graphindex)
    f2, k2 = s.get_graphene()
    mask = np.zeros(c.nxzfl, dtype=bool); mask[u.graphindex] = True
    m18 = np.tile(mask, 18)
    assert np.array_equal(f2[m18], fjn[m18]) and np.array_equal(k2[m18], kfjn[m18])
    assert not f2[~m18].any() and not k2[~m18].any()
    with pytest.raises(NekcemB200Error, match="out of range"):
        s.cem_graphene_current(None, None, u.graphparams, c.yconduc, [c.nxzfl])
    with pytest.raises(NekcemB200Error, match="expected"):
        s.cem_graphene_current(None, None, u.graphparams[:5], c.yconduc, u.graphindex)
    s.close()


def test_rk_tables_default_and_upload():
    """a new context holds rk_storage's LSRK(5,4) tables (src/cem_common.F:86-104), equal to the
    oracle's and -- through the Fortran twin, with the translated reference's COMMON /RKCOEF/ as
    the argument -- to the reference's; param(17) = 22 tables can be uploaded"""
    from oracle import cases
    c = cases.case_boxper((3, 3, 3), 4)
    s = MaxwellB200(3, 4, c.nelt, device=-1)
    a, b, cc = s.get_rk_coefficients()
    assert np.array_equal(a, np.array(c.s.rk4a)) and np.array_equal(b, np.array(c.s.rk4b))
    assert np.array_equal(cc, np.array(c.s.rk4c))
    from oracle import refrun
    if refrun.available():
        r = refrun.ReferenceRun(c)
        L = lib()
        dp = lambda v: v.ctypes.data_as(C.POINTER(C.c_double))
        s.set_rk_coefficients(np.zeros(5), np.zeros(5), np.zeros(6))
        h = C.c_int(s.h)
        L.nekcem_b200_set_rk_coefficients_(C.byref(h), dp(r.view("rk4a")), dp(r.view("rk4b")),
                                           dp(r.view("rk4c")))
        a2, b2, c2 = s.get_rk_coefficients()
        assert np.array_equal(a2, a) and np.array_equal(b2, b) and np.array_equal(c2, cc)
        r.close()
    # rk_storage, ifrk22 branch (:106-110): only the first two entries are set
    s.set_rk_coefficients([0.0, -1.0, 0, 0, 0], [1.0, 0.5, 0, 0, 0], np.zeros(6))
    a3, b3, c3 = s.get_rk_coefficients()
    assert a3[1] == -1.0 and b3[1] == 0.5 and not c3.any()
    s.close()


def test_leading_dimensions_of_fortran_arrays_host_side():
    """SIZE pads lelt, so a .usr's graphene arrays are fjn(lxzfl,3,6), params(lxzfl,12) with
    lxzfl > nxzfl: after nekcem_b200_set_leading_dims the library reads and writes them with that
    leading dimension (host-only context: the packing is host code)."""
    from oracle import cases
    c = cases.case_2dgraphene(1, nx1=4, nel=(3, 6))
    u = c.user
    nf = c.nxzfl
    lf, lp = nf + 3 * c.nxzf * c.nfaces, c.npts + 3 * c.nxyz     # three padding elements
    s = MaxwellB200(2, 4, c.nelt, imode=1, device=-1)
    s.set_faces(c.glo_num, c.cempec[:c.ncempec])
    L = lib()
    assert L.nekcem_b200_set_leading_dims(s.h, c.npts - 1, nf) != 0      # too small: refused
    assert b"smaller" in L.nekcem_b200_last_error()
    assert L.nekcem_b200_set_leading_dims(s.h, lp, lf) == 0
    rng = np.random.default_rng(5)
    pad = lambda a, m: np.concatenate([np.concatenate([a[k * nf:(k + 1) * nf], np.full(lf - nf, np.nan)])
                                       for k in range(m)])
    fjn = rng.standard_normal(18 * nf); kfjn = rng.standard_normal(18 * nf)
    dp = lambda a: a.ctypes.data_as(C.POINTER(C.c_double))
    fp, kp, pp = pad(fjn, 18), pad(kfjn, 18), pad(u.graphparams, 12)
    gi = (u.graphindex + 1).astype(np.int32)
    yc = np.ascontiguousarray(c.yconduc)
    assert L.nekcem_b200_set_graphene(s.h, dp(fp), dp(kp), dp(pp), dp(yc),
                                      gi.ctypes.data_as(C.POINTER(C.c_int32)), gi.size) == 0
    fo = np.zeros(18 * lf); ko = np.zeros(18 * lf)
    assert L.nekcem_b200_get_graphene(s.h, dp(fo), dp(ko)) == 0
    for m in range(18):
        assert np.array_equal(fo[m * lf + u.graphindex], fjn[m * nf + u.graphindex])
        assert np.array_equal(ko[m * lf + u.graphindex], kfjn[m * nf + u.graphindex])
    assert not np.isnan(fo).any()                  # the padding (NaN on input) was never read
    s.close()


def test_twins_report_instead_of_exiting_inside_the_test_runner():
    """tests/conftest.py sets NEKCEM_B200_TWIN_NO_EXIT: a failing Fortran twin prints the library's
    message, counts the error and returns (without the variable it exits, like the reference's
    exitt).  Checked with a twin call that must fail without a device."""
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    L = lib()
    L.nekcem_b200_twin_errors.restype = C.c_int
    before = L.nekcem_b200_twin_errors()
    vals = [C.c_int(v) for v in (3, 4, 27, 3, 1, 0, 0, 0, 0, 1)]
    h = C.c_int(12345)
    L.nekcem_b200_create_(*[C.byref(v) for v in vals], C.byref(h))
    assert L.nekcem_b200_twin_errors() == before + 1 and h.value == -1


def test_empty_registrations_are_accepted_and_remove_the_hook():
    """n = 0 for the incident list and the graphene list (a rank that owns none of the faces
    still makes the call, like the reference's loops over ninc = 0 / ngraph = 0)"""
    b = BoxCase((3, 3, 3), 4)
    s = MaxwellB200(3, 4, b.nelt, device=-1)
    s.set_faces(b.array("glo_num"), b.array("cempec"))
    s.set_incident(np.zeros(0, dtype=np.int64), np.zeros((6, 0)), np.zeros(0), 2.0)
    L = lib()
    z = np.zeros(1)
    dp = z.ctypes.data_as(C.POINTER(C.c_double))
    assert L.nekcem_b200_set_graphene(s.h, None, None, None, None, None, 0) == 0
    assert L.nekcem_b200_get_graphene(s.h, dp, dp) != 0
    assert b"no graphene state" in L.nekcem_b200_last_error()
    s.close()


def test_shim_with_a_padded_size_layout_on_the_host():
    """The Fortran shim driven on the CPU (NEKCEM_B200_HOST_ONLY: the library keeps host copies
    and refuses to compute) with lelt = nelt + 3 as a real SIZE file has it: what
    b200_copy_all_in / b200_update_device hand to the library must be the compact arrays of the
    case -- i.e. the (lpts1,3) leading dimension is honoured for pmlsigma, pmlbn, pmldn, HN, EN
    and the RK registers --, the RK tables are the host's, and b200_update_host writes the
    fields back into the padded COMMON arrays without touching the padding."""
    from oracle import cases, refrun
    if not refrun.available("dropin"):
        pytest.skip("oracle/_ref/libnekcem_ref_dropin.so is not built")
    # nx1 >= 6: the face ids live in COMMON /c_is1/ glo_num(lx1*ly1*lz1*lelv), which holds
    # 6*nx1^2*nelt of them only from nx1 = 6 on
    c = cases.case_3ddielectric(True, nx1=6, nel=(3, 6, 3))
    c.khn[:] = np.linspace(0.0, 1.0, c.khn.size)          # non-trivial RK registers
    os.environ["NEKCEM_B200_HOST_ONLY"] = "1"
    try:
        r = refrun.ReferenceRun(c, kind="dropin", pad_elems=3)
        assert r.lpts1 == c.nxyz * (c.nelt + 3)
        L = lib()
        r.L.b200_copy_all_in_()                            # setup is refused: reported, not fatal
        r.L.b200_update_device_()
        h = int(r.get("b200_handle"))
        dp = lambda a: a.ctypes.data_as(C.POINTER(C.c_double))
        for name in ("pmlsigma", "pmlbn", "pmldn", "hn", "en", "khn", "ken"):
            got = np.zeros(3 * c.npts)
            assert L.nekcem_b200_get_array(h, ARRAY_IDS[name], dp(got), got.size) == 0, name
            assert np.array_equal(got, getattr(c, name)), name
        for name in ("rxmn", "bmn", "hbm1"):
            got = np.zeros(c.npts)
            assert L.nekcem_b200_get_array(h, ARRAY_IDS[name], dp(got), got.size) == 0
            assert np.array_equal(got, getattr(c, name)), name
        got = np.zeros(c.nxzfl)
        assert L.nekcem_b200_get_array(h, ARRAY_IDS["yconduc"], dp(got), got.size) == 0
        a, b, cc = np.zeros(5), np.zeros(5), np.zeros(6)
        assert L.nekcem_b200_get_rk_coefficients(h, dp(a), dp(b), dp(cc)) == 0
        assert np.array_equal(a, np.array(c.s.rk4a)) and np.array_equal(cc, np.array(c.s.rk4c))
        # the way back: poison the COMMON arrays, let the shim fetch the fields
        for k in ("hn", "en"):
            r.view(k)[:] = -7.0
        r.L.b200_update_host_()
        assert np.array_equal(r.field("hn"), c.hn) and np.array_equal(r.field("en"), c.en)
        assert np.all(r.view("hn")[c.npts:r.lpts1] == -7.0)     # padding untouched
        L.nekcem_b200_destroy(h)
        r.close()
    finally:
        del os.environ["NEKCEM_B200_HOST_ONLY"]


def test_shim_registers_graphene_through_the_users_userfsrc_on_the_host():
    """b200_update_device calls the user's userfsrc once on scratch arrays; its
    cem_te_graphene_current call lands in the library twin, which registers the user's arrays
    (here in a host-only context, padded SIZE layout: fjn(lxzfl,3,6) with lxzfl > nxzfl).  The
    twin must not advance or modify the user's arrays."""
    from oracle import cases, refrun
    if not refrun.available("dropin"):
        pytest.skip("oracle/_ref/libnekcem_ref_dropin.so is not built")
    c = cases.case_2dgraphene(1)
    u = c.user
    os.environ["NEKCEM_B200_HOST_ONLY"] = "1"
    try:
        r = refrun.ReferenceRun(c, kind="dropin", pad_elems=3)
        drop, L = refrun.lib("dropin"), lib()
        nf, lf = c.nxzfl, int(r.get("lxzfl"))
        assert lf > nf
        pad = lambda a, m: np.concatenate([np.concatenate([a[k * nf:(k + 1) * nf], np.zeros(lf - nf)])
                                           for k in range(m)])
        fjn, kfjn, resfjn, par = pad(u.fjn, 18), pad(u.kfjn, 18), pad(u.resfjn, 18), pad(u.graphparams, 12)
        f0 = fjn.copy()
        gidx = (u.graphindex + 1).astype(np.int32)
        n = C.c_int(gidx.size)
        dp = lambda a: a.ctypes.data_as(C.POINTER(C.c_double))
        calls = []

        def userfsrc(tt, *src):
            calls.append(tt)
            drop.cem_te_graphene_current_(dp(fjn), dp(kfjn), dp(resfjn), dp(par),
                                          gidx.ctypes.data_as(C.POINTER(C.c_int)), C.byref(n))

        r.put("yconduc", c.yconduc)
        r.set_callback("userfsrc", userfsrc)
        r.L.b200_copy_all_in_()
        r.L.b200_update_device_()
        assert len(calls) == 1 and np.array_equal(fjn, f0)
        h = int(r.get("b200_handle"))
        fo, ko = np.zeros(18 * lf), np.zeros(18 * lf)
        assert L.nekcem_b200_get_graphene(h, dp(fo), dp(ko)) == 0
        for m in range(18):
            assert np.array_equal(fo[m * lf + u.graphindex], u.fjn[m * nf + u.graphindex])
        assert np.abs(fo).max() > 1e-3
        L.nekcem_b200_destroy(h)
        r.close()
    finally:
        del os.environ["NEKCEM_B200_HOST_ONLY"]


def test_shim_registers_drude_arrays_of_a_padded_size_on_the_host():
    """tests/drude with jn(lpts,3), params(lpts,2) dimensioned by a padded SIZE: the .usr's
    usersrc -> cem_maxwell_drude call (made once by b200_update_device) registers them through
    the library twin with the declared leading dimension; get_ade writes them back in the same
    layout (host-only context)."""
    from oracle import cases, refrun
    if not refrun.available("dropin"):
        pytest.skip("oracle/_ref/libnekcem_ref_dropin.so is not built")
    c = cases.case_drude()
    u = c.user
    os.environ["NEKCEM_B200_HOST_ONLY"] = "1"
    try:
        r = refrun.ReferenceRun(c, kind="dropin", pad_elems=3)
        drop, L = refrun.lib("dropin"), lib()
        n, lp = c.npts, r.lpts1
        pad = lambda a, m: np.concatenate([np.concatenate([a[k * n:(k + 1) * n], np.full(lp - n, np.nan)])
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
        h = int(r.get("b200_handle"))
        jo, ko = np.full(3 * lp, -7.0), np.full(3 * lp, -7.0)
        assert L.nekcem_b200_get_ade(h, dp(jo), dp(ko)) == 0
        for k in range(3):
            assert np.array_equal(jo[k * lp:k * lp + n], u.jn[k * n:(k + 1) * n])
            assert np.all(jo[k * lp + n:(k + 1) * lp] == -7.0)      # padding neither read nor written
        assert np.abs(u.jn).max() > 1e-3
        L.nekcem_b200_destroy(h)
        r.close()
    finally:
        del os.environ["NEKCEM_B200_HOST_ONLY"]


def test_f2003_module_interfaces_match_the_header():
    """every `bind(C)` interface of fortran/nekcem_b200_mod.F90 has the argument count (and the
    by-value scalars) of the prototype in include/nekcem_b200.h -- the module is source only (no
    Fortran compiler in the image), so a stale interface would otherwise go unnoticed"""
    import re
    hdr = open(os.path.join(ROOT, "include", "nekcem_b200.h")).read()
    hdr = re.sub(r"/\*.*?\*/", "", hdr, flags=re.S)
    protos = {}
    for m in re.finditer(r"\bint\s+(nekcem_b200_\w+)\s*\(([^;{]*?)\)\s*;", hdr, flags=re.S):
        args = [a.strip() for a in m.group(2).replace("\n", " ").split(",")]
        if args == ["void"] or args == [""]:
            args = []
        protos[m.group(1)] = args
    mod = open(os.path.join(ROOT, "fortran", "nekcem_b200_mod.F90")).read()
    mod = re.sub(r"&\s*\n\s*", " ", mod)
    found = 0
    for m in re.finditer(r"function\s+(nekcem_b200_\w+)\s*\(([^)]*)\)(.*?)end function", mod, flags=re.S | re.I):
        name, fargs, body = m.group(1), [a.strip() for a in m.group(2).split(",") if a.strip()], m.group(3)
        if name == "nekcem_b200_last_error":  # returns const char *, not an int status
            continue
        assert name in protos, f"{name}: in the module but not in the header"
        cargs = protos[name]
        assert len(fargs) == len(cargs), (name, fargs, cargs)
        # scalars passed by value in C must carry the `value` attribute in the interface
        byval, cptr = set(), set()
        for line in body.splitlines():
            if "::" not in line:
                continue
            names = [v.strip().split("(")[0] for v in line.split("::")[1].split(",")]
            if "value" in line.split("::")[0].lower():
                byval.update(names)
            if "c_ptr" in line.split("::")[0].lower():
                cptr.update(names)  # type(c_ptr), value  ==  void * in C
        for fa, ca in zip(fargs, cargs):
            c_is_value = "*" not in ca and "[" not in ca
            assert (fa in byval and fa not in cptr) == c_is_value, (name, fa, ca)
        found += 1
    assert found >= 20
