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


@pytest.mark.parametrize("nx1", [2, 3, 5, 8, 12, 16, 17, 20, 24])
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


# ---------------------------------------------------------------------------------------------
# setup routines whose output the path consumes: also translated from the
# reference, so the oracle's restated setup is pinned too
# ---------------------------------------------------------------------------------------------
@pytest.mark.parametrize("n", list(range(2, 25)))
def test_pin_gll_nodes_weights_and_dgll(n):
    """ZWGLL (-> ZWGLJ -> ZWGLJD -> JACG/JACOBF, ENDW1/2, GAMMAF, PNORMJ) and DGLL (PNLEG),
    src/nek5_speclib.F:107-122, 240-283, 423-521, 807-912"""
    from oracle import oracle as O
    L = refrun.lib()
    z, w = np.zeros(n), np.zeros(n)
    nn = C.c_int(n)
    dp = lambda a: a.ctypes.data_as(C.POINTER(C.c_double))
    L.zwgll_(dp(z), dp(w), C.byref(nn))
    zo, wo = O.zwgll(n)
    assert np.array_equal(z, zo) and np.array_equal(w, wo)
    d, dt = np.zeros(n * n), np.zeros(n * n)
    L.dgll_(dp(d), dp(dt), dp(z), C.byref(nn), C.byref(nn))
    do, dto = O.dgll(zo)
    assert np.array_equal(d, do) and np.array_equal(dt, dto)


def _geometry_case(which):
    from oracle import oracle as O
    if which == "3dboxper":
        return cases.case_3dboxper()
    if which == "3ddielectric":
        return cases.case_3ddielectric(True)
    if which == "drude2d":
        return cases.case_drude()
    if which == "cylwave":
        return cases.case_cylwave(nx1=8)
    # sheared, non-affine 3D elements (curved-metric terms all non-zero)
    mesh = O.box_mesh((2, 2, 2), ((0.0, 1.0),) * 3, ("P  ",) * 6)

    def warp(c):
        x, y, z = c.xm1.copy(), c.ym1.copy(), c.zm1.copy()
        c.xm1[:] = x + 0.05 * np.sin(3 * y) * np.cos(2 * z)
        c.ym1[:] = y + 0.04 * np.sin(2 * x + z)
        c.zm1[:] = z + 0.03 * x * y

    return O.RefCase(mesh, 6, usrdat2=warp)


@pytest.mark.parametrize("which", ["3dboxper", "3ddielectric", "drude2d", "warped", "cylwave"])
def test_pin_geometry(which):
    """GLMAPM1 (XYZRST + cofactors + Jacobian), GEODAT1 (mass bm1 = jac*w3m1) and SETAREA
    (AREA2/AREA3 + UNITVEC): src/nek5_coef.F:555-636, 637-785, 877-925, 992-1237, run on
    the oracle's node coordinates; the cofactors, Jacobian, mass, face areas and normals the
    oracle hands to the path must be the reference's, bit for bit."""
    c = _geometry_case(which)
    r = refrun.ReferenceRun(c)
    n = c.nx1
    nz1 = n if c.ldim == 3 else 1
    for name in ("xm1", "ym1", "zm1"):
        r.put(name, getattr(c, name))
    for name in ("dxm1", "dym1"):
        r.put_opt(name, c.dxm1)
    for name in ("dxtm1", "dytm1"):
        r.put_opt(name, c.dxtm1)
    if c.ldim == 3:
        r.put_opt("dzm1", c.dxm1); r.put_opt("dztm1", c.dxtm1)
    r.put("w3m1", c.w3mn)
    for name in ("wxm1", "wym1"):
        r.put_opt(name, c.wxm1)
    r.put_opt("wzm1", c.wxm1 if c.ldim == 3 else np.ones(1))
    r.put_opt("zgm1", np.concatenate([c.zgm1, c.zgm1, c.zgm1 if c.ldim == 3 else np.zeros(n)]))
    for name in ("rxm1", "rym1", "rzm1", "sxm1", "sym1", "szm1", "txm1", "tym1", "tzm1",
                 "jacm1", "bm1", "area", "unx", "uny", "unz"):
        r.view(name)[:] = -7.0  # poison
    r.L.initds_()
    r.L.glmapm1_()
    r.L.geodat1_()
    for a, b in (("rxmn", "rxm1"), ("rymn", "rym1"), ("sxmn", "sxm1"), ("symn", "sym1"),
                 ("jacm", "jacm1"), ("bmn", "bm1")) + ((("rzmn", "rzm1"), ("szmn", "szm1"),
                 ("txmn", "txm1"), ("tymn", "tym1"), ("tzmn", "tzm1")) if c.ldim == 3 else ()):
        x, y = getattr(c, a), r.view(b)[:c.npts]
        assert np.array_equal(x, y), (a, float(np.abs(x - y).max()))
    for a, b in (("aream", "area"), ("unxm", "unx"), ("unym", "uny")) + (
            (("unzm", "unz"),) if c.ldim == 3 else ()):
        x, y = getattr(c, a), r.view(b)[:c.nxzfl]
        assert np.array_equal(x, y), (a, float(np.abs(x - y).max()))
    r.close()


def _dp(a):
    return a.ctypes.data_as(C.POINTER(C.c_double))


@pytest.mark.parametrize("which", ["3dboxpec", "3ddielectric", "drude2d", "cylwave"])
def test_pin_materials_and_pec_list(which):
    """cem_maxwell_materials (impedances Y_0,Y_1,Z_0,Z_1 via the reference's gs_op, PEC
    doubling, src/cem_maxwell.F:262-325) and cem_maxwell_pec_init (:1338-1366)"""
    c = {"3dboxpec": cases.case_3dboxpec, "3ddielectric": lambda: cases.case_3ddielectric(True),
         "drude2d": cases.case_drude, "cylwave": lambda: cases.case_cylwave(nx1=6)}[which]()
    r = refrun.ReferenceRun(c)
    r.set_cbc(c.mesh.cbc)
    for name in ("y_0", "y_1", "z_0", "z_1"):
        r.view(name)[:] = 0.0
    r.L.cem_maxwell_materials_()
    for name in ("Y_0", "Y_1", "Z_0", "Z_1"):
        assert np.array_equal(getattr(c, name), r.view(name.lower())[:c.nxzfl]), name
    r.view("cempec")[:] = 0
    r.L.cem_maxwell_pec_init_()
    assert r.get("ncempec") == c.ncempec
    assert np.array_equal(r.view("cempec")[:c.ncempec], c.cempec[:c.ncempec] + 1)
    r.close()


@pytest.mark.parametrize("which", ["3ddielectric", "3dboxpml", "drude2d"])
def test_pin_pml_setup(which):
    """pml_fill_faceary / march_faces (gs_op max) / pml_extent_and_tags / pml_calc_sigma,
    src/cem_maxwell_pml.F:85-506: PML element list, tags, extents and the sigma profile"""
    c = {"3ddielectric": lambda: cases.case_3ddielectric(True),
         "3dboxpml": lambda: cases.case_3dboxpml(nx1=6, nel=(5, 5, 5)),
         "drude2d": cases.case_drude}[which]()
    r = refrun.ReferenceRun(c)
    r.set_cbc(c.mesh.cbc)
    for name in ("xm1", "ym1", "zm1"):
        r.put(name, getattr(c, name))
    for a, b in (("rxmn", "rxm1"), ("rymn", "rym1"), ("rzmn", "rzm1"), ("sxmn", "sxm1"),
                 ("symn", "sym1"), ("szmn", "szm1"), ("txmn", "txm1"), ("tymn", "tym1"),
                 ("tzmn", "tzm1")):
        r.put_opt(b, getattr(c, a))
    # the reference passes its COMMON arrays as the actual arguments (src/cem_maxwell.F:164-175)
    faceary = np.zeros(c.nxzfl)  # /scratch/ faceary: not named by any translated routine
    tag = r.view("pmltag")
    inner, outer = np.zeros(2 * c.ldim), np.zeros(2 * c.ldim)
    ip = lambda a: a.ctypes.data_as(C.POINTER(C.c_int))
    thick = C.c_int(c.pmlthick)
    r.L.pml_fill_faceary_(_dp(faceary), C.byref(thick))
    r.L.pml_extent_and_tags_(_dp(inner), _dp(outer), ip(tag), _dp(faceary))
    assert np.array_equal(tag[:c.nelt], c.pmltag)
    assert np.array_equal(inner[:2 * c.ldim], c.pmlinner)
    assert np.array_equal(outer[:2 * c.ldim], c.pmlouter)
    r.view("pmlsigma")[:] = 0.0
    order, referr = C.c_double(c.pmlorder), C.c_double(c.pmlreferr)
    r.L.pml_calc_sigma_(_dp(inner), _dp(outer), ip(tag), C.byref(order), C.byref(referr))
    assert r.get("maxpml") == c.maxpml
    assert np.array_equal(r.view("pmlptr")[:c.maxpml], c.pmlptr[:c.maxpml] + 1)
    assert np.array_equal(r.view("pmlsigma")[:3 * c.npts], c.pmlsigma)
    r.close()


def test_pin_error_norms_and_dxmin():
    """cem_error (src/cem_common.F:1335-1355: the L2 / Linf norms userchk prints) and get_dxmin
    (src/nek5_courant.F:2-79: the CFL length of set_dt)"""
    from oracle import oracle as O
    c = cases.case_3dboxper()
    r = refrun.ReferenceRun(c)
    c.step(3)
    shn, sen = c.usersol(c, c.time)
    r.set("volvm1", c.volvm1)
    n = C.c_int(c.npts)
    err = np.zeros(c.npts)
    for arr, sol in ((c.hn, shn), (c.en, sen)):
        for k in range(3):
            u = np.ascontiguousarray(c.comp(arr, k)); ex = np.ascontiguousarray(c.comp(sol, k))
            l2, linf = C.c_double(), C.c_double()
            r.L.cem_error_(_dp(u), _dp(ex), _dp(err), C.byref(n), C.byref(l2), C.byref(linf))
            a, b = c.cem_error(u, ex)
            assert (a, b) == (l2.value, linf.value)
            assert l2.value > 0
    for name in ("xm1", "ym1", "zm1"):
        r.put(name, getattr(c, name))
    d = C.c_double()
    r.L.get_dxmin_(C.byref(d))
    assert d.value == c.L.ora_get_dxmin(c.ldim, c.nx1, c.nelt, O.dp(c.xm1), O.dp(c.ym1), O.dp(c.zm1))
    r.close()


@pytest.mark.parametrize("which", ["3dboxper", "3dboxpec", "2dboxper-te", "2dboxper-tm",
                                   "2dboxpec-te", "2dboxpec-tm"])
def test_pin_usersol_and_error_norms(which):
    """The analytic solutions compiled into the reference's .usr files (usersol of
    tests/3dboxper, 3dboxpec, 2dboxper, 2dboxpec), translated from those files: the oracle's
    numpy restatements agree to round-off of the libm / numpy sin and cos (<= 4 ulp of an O(1)
    value), and the L2 / Linf errors the reference's userchk would print for the oracle's
    fields -- reference usersol + reference cem_error -- equal the oracle's own."""
    name, _, mode = which.partition("-")
    imode = {"te": 1, "tm": 2}.get(mode, 3)
    c = {"3dboxper": cases.case_3dboxper, "3dboxpec": cases.case_3dboxpec,
         "2dboxper": lambda: cases.case_2dboxper(imode),
         "2dboxpec": lambda: cases.case_2dboxpec(imode)}[name]()
    r = refrun.ReferenceRun(c)
    for nm in ("xm1", "ym1", "zm1"):
        r.put_opt(nm, getattr(c, nm))
    c.step(7)
    n = c.npts
    sol = [np.zeros(n) for _ in range(6)]
    tt = C.c_double(c.time)
    getattr(r.L, "usersol__%s_" % name)(C.byref(tt), *[_dp(a) for a in sol])
    shn, sen = c.usersol(c, c.time)
    mine = [c.comp(shn, k) for k in range(3)] + [c.comp(sen, k) for k in range(3)]
    for k in range(6):
        assert np.max(np.abs(sol[k] - mine[k])) <= 1e-15, (k, np.max(np.abs(sol[k] - mine[k])))
    assert max(np.abs(a).max() for a in sol) > 0.1
    # userchk: cem_error of every component against the reference's usersol
    r.set("volvm1", c.volvm1)
    l2o, linfo = c.errors(c.usersol)
    err = np.zeros(n)
    nn = C.c_int(n)
    for k in range(6):
        u = np.ascontiguousarray(c.comp(c.hn if k < 3 else c.en, k % 3))
        l2, linf = C.c_double(), C.c_double()
        r.L.cem_error_(_dp(u), _dp(sol[k]), _dp(err), C.byref(nn), C.byref(l2), C.byref(linf))
        assert l2.value <= c.tol["l2"][k] and linf.value <= c.tol["linf"][k]
        assert abs(l2.value - l2o[k]) <= 1e-3 * l2o[k] + 1e-15
        assert abs(linf.value - linfo[k]) <= 1e-3 * linfo[k] + 1e-14
    r.close()


def test_pin_rotated_element_frames():
    """elements with arbitrarily oriented local frames (all 24 proper rotations): the reference's
    gs_op_fields pairs the face points through the ids alone; oracle == translated reference"""
    c, rots = cases.case_boxper_rotated((3, 3, 3), 5, dt=-1e-3)
    r = refrun.ReferenceRun(c)
    c.step(5); r.step(5)
    _assert_same(c, r)
    r.close()


@pytest.mark.parametrize("which", ["3dboxper", "drude2d", "box3d"])
def test_pin_genxyz(which):
    """GENXYZ (src/nek5_genxyz.F:562-680): GLL node coordinates of every element from its
    corner vertices xc,yc,zc (trilinear blend in the reference's summation order).  The
    oracle's coordinates BEFORE the user's usrdat2 rescaling must be the reference's."""
    from oracle import oracle as O
    if which == "3dboxper":
        mesh, _ = cases._load_mesh("3dboxper_mesh.npz")     # vertices of the reference's .re2
        nx1 = 9
    elif which == "drude2d":
        mesh = O.box_mesh((4, 32), ((-1500.0, 1500.0),) * 2, ("P  ", "P  ", "PEC", "PML"))
        nx1 = 9
    else:
        mesh = O.box_mesh((3, 2, 4), ((0.0, 1.0), (-2.0, 5.0), (0.5, 0.75)), ("P  ",) * 6,
                          gain=(1.0, 1.3, 0.8))
        nx1 = 6
    c = O.RefCase(mesh, nx1, imode=1)                        # no usrdat2: raw genxyz output
    r = refrun.ReferenceRun(c)
    ldim = c.ldim
    nc = 2 ** ldim
    # xc(8,lelt), yc, zc in preprocessor corner order (src/INPUT); zgm1 = GLL points per axis
    for name, arr in (("xc", mesh.xc), ("yc", mesh.yc), ("zc", mesh.zc)):
        v = r.view(name).reshape(-1)
        v[:] = 0.0
        full = np.zeros((c.nelt, 8))
        full[:, :nc] = arr[:, :nc]
        v[:8 * c.nelt] = full.reshape(-1)
    n = nx1
    r.put("zgm1", np.concatenate([c.zgm1, c.zgm1, c.zgm1 if ldim == 3 else np.zeros(n)]))
    for k in ("ifgmsh3", "ifaxis"):
        try:
            r.set(k, 0)
        except KeyError:
            pass
    try:
        r.view_char("ccurve")[:] = ord(" ")
    except KeyError:
        pass
    x, y, z = np.zeros(c.npts), np.zeros(c.npts), np.zeros(c.npts)
    nz1 = nx1 if ldim == 3 else 1
    r.L.genxyz_(_dp(x), _dp(y), _dp(z), C.byref(C.c_int(nx1)), C.byref(C.c_int(nx1)),
                C.byref(C.c_int(nz1)))
    assert np.array_equal(x, c.xm1) and np.array_equal(y, c.ym1)
    if ldim == 3:
        assert np.array_equal(z, c.zm1)
    r.close()


# ---------------------------------------------------------------------------------------------
# the shipped .usr files themselves (translated from /root/reference/tests/<case>/<case>.usr):
# pins the numpy restatements of the user callbacks in oracle/cases.py
# ---------------------------------------------------------------------------------------------
def _usr(L, name, case):
    return getattr(L, "%s__%s_" % (name, case))


@pytest.mark.parametrize("case", ["3dboxper", "3dboxpec", "3ddielectric", "3dboxpml"])
def test_pin_usrdat2_mesh_rescaling(case):
    """usrdat2 of the .usr files: the affine rescaling of the mesh (glmin/glmax of the
    coordinates, then the per-node formula) applied to the raw genxyz coordinates"""
    from oracle import oracle as O
    build = {"3dboxper": cases.case_3dboxper, "3dboxpec": cases.case_3dboxpec,
             "3ddielectric": lambda: cases.case_3ddielectric(False),
             "3dboxpml": lambda: cases.case_3dboxpml(nx1=5, nel=(4, 4, 4))}[case]
    c = build()                                      # coordinates AFTER the oracle's usrdat2
    raw = O.RefCase(c.mesh, c.nx1)                   # raw genxyz output
    r = refrun.ReferenceRun(c)
    for nm in ("xm1", "ym1", "zm1"):
        r.put(nm, getattr(raw, nm))
    _usr(r.L, "usrdat2", case)()
    for nm in ("xm1", "ym1", "zm1"):
        assert np.array_equal(r.view(nm)[:c.npts], getattr(c, nm)), nm
    r.close()


@pytest.mark.parametrize("twomat", [False, True])
def test_pin_3ddielectric_with_the_shipped_usr(twomat):
    """tests/3ddielectric driven by ITS OWN .usr: uservp (materials + the incident-face index),
    userini -> usersol (initial fields incl. the PML decay factor, pmlbn/pmldn), userinc in every
    stage.  Materials and the index must equal the oracle's restatement exactly; fields agree to
    the round-off of libm vs numpy cos/exp (<= 1e-13 relative after 10 steps)."""
    c = cases.case_3ddielectric(twomat)
    r = refrun.ReferenceRun(c)
    L = r.L
    r.set_cbc(c.mesh.cbc)
    for nm in ("xm1", "ym1", "zm1"):
        r.put(nm, getattr(c, nm))
    r.view("param")[69] = 1.0 if twomat else 0.0       # param(70)
    # PML layout first (cem_maxwell_init order, src/cem_maxwell.F:164-175), all reference code
    for a, b in (("rxmn", "rxm1"), ("rymn", "rym1"), ("rzmn", "rzm1"), ("sxmn", "sxm1"),
                 ("symn", "sym1"), ("szmn", "szm1"), ("txmn", "txm1"), ("tymn", "tym1"),
                 ("tzmn", "tzm1")):
        r.put_opt(b, getattr(c, a))
    faceary = np.zeros(c.nxzfl)
    ip = lambda a: a.ctypes.data_as(C.POINTER(C.c_int))
    L.pml_fill_faceary_(_dp(faceary), C.byref(C.c_int(c.pmlthick)))
    L.pml_extent_and_tags_(_dp(r.view("pmlinner")), _dp(r.view("pmlouter")), ip(r.view("pmltag")),
                           _dp(faceary))
    one = C.c_int(1)
    r.view("permittivity")[:] = 0.0
    _usr(L, "uservp", "3ddielectric")(C.byref(one), C.byref(one), C.byref(one), C.byref(one))
    assert np.array_equal(r.view("permittivity")[:c.npts], c.permittivity)
    assert np.array_equal(r.view("permeability")[:c.npts], c.permeability)
    ninc = int(r.get("ninc__3ddielectric"))
    assert ninc == c.user.incindex.size
    assert np.array_equal(r.view("incindex__3ddielectric")[:ninc], c.user.incindex + 1)
    L.pml_calc_sigma_(_dp(r.view("pmlinner")), _dp(r.view("pmlouter")), ip(r.view("pmltag")),
                      C.byref(C.c_double(c.pmlorder)), C.byref(C.c_double(c.pmlreferr)))
    r.set("pmlorder", c.pmlorder); r.set("pmlreferr", c.pmlreferr)
    # userini -> usersol: initial condition
    n, n3 = c.npts, 3 * c.npts
    hn, en = r.view("hn"), r.view("en")
    hn[:] = 0.0; en[:] = 0.0
    tt = C.c_double(0.0)
    _usr(L, "userini", "3ddielectric")(C.byref(tt), _dp(hn[0:]), _dp(hn[n:]), _dp(hn[2 * n:]),
                                      _dp(en[0:]), _dp(en[n:]), _dp(en[2 * n:]))
    shn, sen = c.user.usersol(c, 0.0)
    assert np.abs(hn[:n3] - shn).max() <= 2e-15 and np.abs(en[:n3] - sen).max() <= 2e-15
    assert np.abs(r.view("pmldn")[:n3] - c.pmldn).max() <= 4e-15
    # time stepping with the .usr's userinc as the callback
    L.ref_set_user(0, C.cast(_usr(L, "userinc", "3ddielectric"), refrun.USERCB))
    c.step(10); r.step(10)
    num = np.sqrt(np.sum((c.hn - hn[:n3]) ** 2) + np.sum((c.en - en[:n3]) ** 2))
    den = np.sqrt(np.sum(c.hn ** 2) + np.sum(c.en ** 2))
    assert num / den <= 1e-13
    # and the .usr's usersol at the end time: the reference's userchk tolerances hold
    sol = [np.zeros(n) for _ in range(6)]
    tt = C.c_double(c.time)
    _usr(L, "usersol", "3ddielectric")(C.byref(tt), *[_dp(a) for a in sol])
    mine_h, mine_e = c.user.usersol(c, c.time)
    for k in range(3):
        assert np.abs(sol[k] - c.comp(mine_h, k)).max() <= 2e-15
        assert np.abs(sol[3 + k] - c.comp(mine_e, k)).max() <= 2e-15
    r.close()


def test_pin_3dboxpml_with_the_shipped_usersrc():
    """tests/3dboxpml driven by its own usersrc (Gaussian dipole, 3dboxpml.usr:30-88)"""
    c = cases.case_3dboxpml(nx1=6, nel=(5, 5, 5))
    r = refrun.ReferenceRun(c)
    for nm in ("xm1", "ym1", "zm1"):
        r.put(nm, getattr(c, nm))
    r.L.ref_set_user(1, C.cast(_usr(r.L, "usersrc", "3dboxpml"), refrun.USERCB))
    c.step(15); r.step(15)
    n3 = 3 * c.npts
    num = np.sqrt(np.sum((c.hn - r.hn) ** 2) + np.sum((c.en - r.en) ** 2))
    den = np.sqrt(np.sum(c.hn ** 2) + np.sum(c.en ** 2))
    assert den > 1e-8 and num / den <= 1e-13
    r.close()


def test_pin_drude_with_the_shipped_userinc_and_usersrc():
    """tests/drude driven by its own userinc (plane-wave injection) and usersrc (which calls the
    reference's cem_maxwell_drude on the arrays of the .usr's COMMON /userdrude/).  The analytic
    solution of this case is evaluated in COMPLEX arithmetic by the .usr and is not translated;
    the initial state and the material tables come from the oracle's restatement."""
    c = cases.case_drude()
    u = c.user
    r = refrun.ReferenceRun(c)
    L = r.L
    r.put("ym1", c.ym1)
    n = c.npts
    # COMMON /userparam/ and /userincvars/ (drude.usr:20-30), /userdrude/ (:53-56)
    # (the .usr's own COMMON blocks get per-file globals: <name>__drude)
    r.set("omega__drude", u.omega); r.set("k1__drude", u.k1); r.set("eta1__drude", u.eta1)
    r.put("incindex__drude", u.incindex + 1); r.set("ninc__drude", u.incindex.size)
    r.put("jn__drude", u.jn); r.put("kjn__drude", u.kjn)
    r.put("drudeparams__drude", u.params)
    r.put("drudeindex__drude", u.index + 1); r.set("ndrude__drude", u.index.size)
    L.ref_set_user(0, C.cast(_usr(L, "userinc", "drude"), refrun.USERCB))
    L.ref_set_user(1, C.cast(_usr(L, "usersrc", "drude"), refrun.USERCB))
    c.step(40); r.step(40)
    num = np.sqrt(np.sum((c.hn - r.hn) ** 2) + np.sum((c.en - r.en) ** 2))
    den = np.sqrt(np.sum(c.hn ** 2) + np.sum(c.en ** 2))
    assert num / den <= 1e-13
    j = r.view("jn__drude")[:3 * n]
    assert np.abs(j).max() > 1e-3
    assert np.sqrt(np.sum((j - u.jn) ** 2)) / np.sqrt(np.sum(u.jn ** 2)) <= 1e-13
    r.close()


@pytest.mark.parametrize("kind", ["drude", "lorentz"])
def test_pin_dispersive_case_with_its_whole_usr(kind):
    """tests/drude and tests/lorentz driven entirely by their own .usr (complex arithmetic incl.
    the csqrt branch choices, translated to C99 double _Complex): uservp (materials, ADE
    parameter table, node list, incident-face index), userini -> usersol (initial fields, PML
    fields, initial currents), userinc + usersrc in every stage, usersol at the end.  The integer
    tables must equal the oracle's restatement exactly, the real ones to round-off."""
    c = cases.case_drude() if kind == "drude" else cases.case_lorentz()
    u = c.user
    sfx = "__" + kind
    r = refrun.ReferenceRun(c)
    L = r.L
    r.put("ym1", c.ym1); r.put("xm1", c.xm1)
    # PML description the .usr's usersol reads (src/PML)
    r.put("pmltag", c.pmltag); r.put("pmlinner", c.pmlinner); r.put("pmlouter", c.pmlouter)
    r.set("pmlorder", c.pmlorder); r.set("pmlreferr", c.pmlreferr)
    n, n3 = c.npts, 3 * c.npts
    one = C.c_int(1)
    r.view("permittivity")[:] = 0.0
    _usr(L, "uservp", kind)(C.byref(one), C.byref(one), C.byref(one), C.byref(one))
    assert np.array_equal(r.view("permittivity")[:n], c.permittivity)
    assert np.array_equal(r.view("permeability")[:n], c.permeability)
    idx_name, n_name, par_name = (("drudeindex", "ndrude", "drudeparams") if kind == "drude"
                                  else ("lorentzindex", "nlorentz", "lorentzparams"))
    nade = int(r.get(n_name + sfx))
    assert nade == u.index.size
    assert np.array_equal(r.view(idx_name + sfx)[:nade], u.index + 1)
    assert np.array_equal(r.view(par_name + sfx)[:u.params.size], u.params)
    ninc = int(r.get("ninc" + sfx))
    assert ninc == u.incindex.size and np.array_equal(r.view("incindex" + sfx)[:ninc], u.incindex + 1)
    # the complex material constants: same branch of the square roots
    assert abs(r.view("eta2" + sfx)[0] - u.eta2) <= 1e-15 * abs(u.eta2)
    assert abs(r.view("k2" + sfx)[0] - u.k2) <= 1e-15 * abs(u.k2)
    assert abs(r.view("tran" + sfx)[0] - u.tran) <= 1e-15 * abs(u.tran)
    # userini -> usersol
    hn, en = r.view("hn"), r.view("en")
    c0 = cases.case_drude() if kind == "drude" else cases.case_lorentz()   # pristine t=0 state
    hn[:] = 0.0; en[:] = 0.0
    tt = C.c_double(0.0)
    _usr(L, "userini", kind)(C.byref(tt), _dp(hn[0:]), _dp(hn[n:]), _dp(hn[2 * n:]),
                             _dp(en[0:]), _dp(en[n:]), _dp(en[2 * n:]))
    scale = max(np.abs(c0.hn).max(), np.abs(c0.en).max())
    assert np.abs(hn[:n3] - c0.hn).max() <= 1e-14 * scale
    assert np.abs(en[:n3] - c0.en).max() <= 1e-14 * scale
    j0 = r.view("jn" + sfx)[:c0.user.jn.size]
    assert np.abs(j0 - c0.user.jn).max() <= 1e-13 * np.abs(c0.user.jn).max()
    assert np.abs(r.view("pmldn")[:n3] - c0.pmldn).max() <= 1e-14 * scale
    # time stepping with the .usr's userinc and usersrc
    L.ref_set_user(0, C.cast(_usr(L, "userinc", kind), refrun.USERCB))
    L.ref_set_user(1, C.cast(_usr(L, "usersrc", kind), refrun.USERCB))
    c.step(40); r.step(40)
    num = np.sqrt(np.sum((c.hn - hn[:n3]) ** 2) + np.sum((c.en - en[:n3]) ** 2))
    den = np.sqrt(np.sum(c.hn ** 2) + np.sum(c.en ** 2))
    assert num / den <= 1e-12
    # the .usr's usersol at the end time agrees with the restatement; userchk tolerances hold
    sol = [np.zeros(n) for _ in range(6)]
    tt = C.c_double(c.time)
    _usr(L, "usersol", kind)(C.byref(tt), *[_dp(a) for a in sol])
    mh, me = u.usersol(c, c.time)
    for k in range(3):
        assert np.abs(sol[k] - c.comp(mh, k)).max() <= 1e-14 * scale
        assert np.abs(sol[3 + k] - c.comp(me, k)).max() <= 1e-14 * scale
    r.set("volvm1", c.volvm1)
    err = np.zeros(n)
    nn = C.c_int(n)
    for k in (2, 3, 4):                                  # hz, ex, ey (drude.usr userchk)
        fld = np.ascontiguousarray(c.comp(c.hn if k < 3 else c.en, k % 3))
        l2, linf = C.c_double(), C.c_double()
        L.cem_error_(_dp(fld), _dp(sol[k]), _dp(err), C.byref(nn), C.byref(l2), C.byref(linf))
        assert l2.value <= c.tol["l2"][k] and linf.value <= c.tol["linf"][k]
    r.close()


# ---------------------------------------------------------------------------------------------
# graphene sheets (SURVEY.md 8f rank 1 userfsrc / rank 4 surface-current ADEs):
# cem_3d/te/tm_graphene_current (src/cem_maxwell.F:2827-3093) and the .usr files that call them
# ---------------------------------------------------------------------------------------------
def _graphene_case(which, small=True):
    if which == "3dgraphene":
        return cases.case_3dgraphene(nel=(3, 12, 3) if small else (4, 12, 4))
    return cases.case_2dgraphene(1 if which.endswith("te") else 2)


_GRAPHENE = ["3dgraphene", "2dgraphene-te", "2dgraphene-tm"]


@pytest.mark.parametrize("which", _GRAPHENE)
def test_pin_graphene_current(which):
    """the oracle's restatement of the three graphene-current routines equals the translated
    reference bit for bit: fields, RK registers, PML state, and the sheet currents fjn/kfjn;
    userfsrc is a test-side callback calling the reference's routine on its own arrays"""
    c = _graphene_case(which)
    u = c.user
    r = refrun.ReferenceRun(c)
    r.put("yconduc", c.yconduc)
    r.set_callback("userinc", u.userinc(c))
    fjn, kfjn, resfjn = u.fjn.copy(), u.kfjn.copy(), u.resfjn.copy()
    params = u.graphparams.copy()
    gidx = (u.graphindex + 1).astype(np.int32)
    n = C.c_int(gidx.size)
    fn = {3: r.L.cem_3d_graphene_current_, 1: r.L.cem_te_graphene_current_,
          2: r.L.cem_tm_graphene_current_}[c.imode]
    dp = lambda a: a.ctypes.data_as(C.POINTER(C.c_double))
    nf = c.nxzfl
    comps = {3: (0, 1, 2), 1: (0, 1), 2: (2,)}[c.imode]
    j = u.graphindex

    def userfsrc(tt, shx, shy, shz, sex, sey, sez):  # 3dgraphene.usr:236-266
        fn(dp(fjn), dp(kfjn), dp(resfjn), dp(params), gidx.ctypes.data_as(C.POINTER(C.c_int)),
           C.byref(n))
        src = (shx, shy, shz)
        for q in comps:
            src[q][j] = src[q][j] - fjn[q * nf + j]

    r.set_callback("userfsrc", userfsrc)
    c.step(20); r.step(20)
    _assert_same(c, r)
    assert np.array_equal(u.fjn, fjn) and np.array_equal(u.kfjn, kfjn)
    assert np.abs(fjn).max() > 1e-3 and np.abs(kfjn).max() > 0
    r.close()


@pytest.mark.parametrize("which", _GRAPHENE)
def test_pin_graphene_case_with_its_whole_usr(which):
    """tests/3dgraphene and tests/2dgraphene (TE, TM) driven entirely by their own .usr: usrdat2,
    uservp (materials, complex sheet conductivity, reflection / transmission coefficients,
    incident-face and graphene-face index, parameter table), userini -> usersol (fields, PML
    fields, initial sheet currents), userinc + userfsrc in every stage, usersol at the end and
    the userchk tolerances evaluated by the reference's cem_error."""
    c = _graphene_case(which)
    u = c.user
    kind = which.split("-")[0]
    sfx = "__" + kind
    r = refrun.ReferenceRun(c)
    L = r.L
    r.put("xm1", c.xm1); r.put("ym1", c.ym1)
    if c.ldim == 3:
        r.put("zm1", c.zm1)
    r.put("yconduc", c.yconduc)
    r.put("pmltag", c.pmltag); r.put("pmlinner", c.pmlinner); r.put("pmlouter", c.pmlouter)
    r.set("pmlorder", c.pmlorder); r.set("pmlreferr", c.pmlreferr)
    try:
        r.set("if3d", int(c.ldim == 3))
    except KeyError:
        pass
    n, n3, nf = c.npts, 3 * c.npts, c.nxzfl
    one = C.c_int(1)
    r.view("permittivity")[:] = 0.0
    _usr(L, "uservp", kind)(C.byref(one), C.byref(one), C.byref(one), C.byref(one))
    assert np.array_equal(r.view("permittivity")[:n], c.permittivity)
    assert np.array_equal(r.view("permeability")[:n], c.permeability)
    ng = int(r.get("ngraph" + sfx))
    assert ng == u.graphindex.size
    assert np.array_equal(r.view("graphindex" + sfx)[:ng], u.graphindex + 1)
    assert np.array_equal(r.view("graphparams" + sfx)[:12 * nf], u.graphparams)
    ninc = int(r.get("ninc" + sfx))
    assert ninc == u.incindex.size and np.array_equal(r.view("incindex" + sfx)[:ninc], u.incindex + 1)
    assert abs(r.view("sigmagraph" + sfx)[0] - u.sigmagraph) <= 4e-16 * abs(u.sigmagraph)
    if kind == "3dgraphene":
        coefs = (("reflte", u.reflte), ("trante", u.trante), ("refltm", u.refltm), ("trantm", u.trantm))
    else:
        coefs = ((("refl", u.reflte), ("tran", u.trante)) if c.imode == 1
                 else (("refl", u.refltm), ("tran", u.trantm)))
    for name, val in coefs:
        assert abs(r.view(name + sfx)[0] - val) <= 4e-16 * abs(val), name
    # userini -> usersol
    hn, en = r.view("hn"), r.view("en")
    c0 = _graphene_case(which)  # pristine t=0 state
    hn[:] = 0.0; en[:] = 0.0
    tt = C.c_double(0.0)
    _usr(L, "userini", kind)(C.byref(tt), _dp(hn[0:]), _dp(hn[n:]), _dp(hn[2 * n:]),
                             _dp(en[0:]), _dp(en[n:]), _dp(en[2 * n:]))
    scale = max(np.abs(c0.hn).max(), np.abs(c0.en).max())
    assert np.abs(hn[:n3] - c0.hn).max() <= 1e-14 * scale
    assert np.abs(en[:n3] - c0.en).max() <= 1e-14 * scale
    f0 = r.view("fjn" + sfx)[:18 * nf]
    assert np.abs(f0 - c0.user.fjn).max() <= 1e-13 * np.abs(c0.user.fjn).max()
    assert np.abs(r.view("pmldn")[:n3] - c0.pmldn).max() <= 1e-14 * scale
    assert np.abs(r.view("pmlbn")[:n3] - c0.pmlbn).max() <= 1e-14 * scale
    # time stepping with the .usr's userinc and userfsrc
    L.ref_set_user(0, C.cast(_usr(L, "userinc", kind), refrun.USERCB))
    L.ref_set_user(2, C.cast(_usr(L, "userfsrc", kind), refrun.USERCB))
    c.step(40); r.step(40)
    num = np.sqrt(np.sum((c.hn - hn[:n3]) ** 2) + np.sum((c.en - en[:n3]) ** 2))
    den = np.sqrt(np.sum(c.hn ** 2) + np.sum(c.en ** 2))
    assert num / den <= 1e-12
    fj = r.view("fjn" + sfx)[:18 * nf]
    assert np.sqrt(np.sum((fj - u.fjn) ** 2)) <= 1e-12 * np.sqrt(np.sum(u.fjn ** 2))
    # the .usr's usersol at the end time agrees with the restatement; userchk tolerances hold
    sol = [np.zeros(n) for _ in range(6)]
    tt = C.c_double(c.time)
    _usr(L, "usersol", kind)(C.byref(tt), *[_dp(a) for a in sol])
    mh, me = u.usersol(c, c.time)
    for k in range(3):
        assert np.abs(sol[k] - c.comp(mh, k)).max() <= 1e-14 * scale
        assert np.abs(sol[3 + k] - c.comp(me, k)).max() <= 1e-14 * scale
    r.set("volvm1", c.volvm1)
    err = np.zeros(n)
    nn = C.c_int(n)
    for k in range(6):
        if c.tol["l2"][k] == 0:
            continue
        fld = np.ascontiguousarray(c.comp(c.hn if k < 3 else c.en, k % 3))
        l2, linf = C.c_double(), C.c_double()
        L.cem_error_(_dp(fld), _dp(sol[k]), _dp(err), C.byref(nn), C.byref(l2), C.byref(linf))
        assert l2.value <= c.tol["l2"][k] and linf.value <= c.tol["linf"][k], (k, l2.value, linf.value)
    r.close()


# ---------------------------------------------------------------------------------------------
# the remaining 2D Maxwell configurations of the reference's test suite
# ---------------------------------------------------------------------------------------------
@pytest.mark.parametrize("imode", [1, 2])
@pytest.mark.parametrize("twomat", [False, True])
def test_pin_2ddielectric_with_its_whole_usr(imode, twomat):
    """tests/2ddielectric (TE/TM x one/two materials) driven by its own .usr: usrdat2, uservp
    (materials from param(70), reflection/transmission coefficient of the mode, incident-face
    index), userini -> usersol (fields incl. the PML decay factor, pmlbn/pmldn), userinc in every
    stage, usersol at the end and the userchk tolerances evaluated by the reference's cem_error."""
    from oracle import oracle as O
    c = cases.case_2ddielectric(imode, twomat)
    u = c.user
    sfx = "__2ddielectric"
    r = refrun.ReferenceRun(c)
    L = r.L
    raw = O.RefCase(c.mesh, c.nx1, imode=imode)
    r.put("xm1", raw.xm1); r.put("ym1", raw.ym1)
    _usr(L, "usrdat2", "2ddielectric")()
    assert np.array_equal(r.view("xm1")[:c.npts], c.xm1)
    assert np.array_equal(r.view("ym1")[:c.npts], c.ym1)
    r.view("param")[69] = 1.0 if twomat else 0.0       # param(70)
    r.put("pmltag", c.pmltag); r.put("pmlinner", c.pmlinner); r.put("pmlouter", c.pmlouter)
    r.set("pmlorder", c.pmlorder); r.set("pmlreferr", c.pmlreferr)
    n, n3 = c.npts, 3 * c.npts
    one = C.c_int(1)
    r.view("permittivity")[:] = 0.0
    _usr(L, "uservp", "2ddielectric")(C.byref(one), C.byref(one), C.byref(one), C.byref(one))
    assert np.array_equal(r.view("permittivity")[:n], c.permittivity)
    assert np.array_equal(r.view("permeability")[:n], c.permeability)
    ninc = int(r.get("ninc" + sfx))
    assert ninc == u.incindex.size and np.array_equal(r.view("incindex" + sfx)[:ninc], u.incindex + 1)
    assert r.get("refl" + sfx) == u.refl and r.get("tran" + sfx) == u.tran
    hn, en = r.view("hn"), r.view("en")
    hn[:] = 0.0; en[:] = 0.0
    tt = C.c_double(0.0)
    _usr(L, "userini", "2ddielectric")(C.byref(tt), _dp(hn[0:]), _dp(hn[n:]), _dp(hn[2 * n:]),
                                      _dp(en[0:]), _dp(en[n:]), _dp(en[2 * n:]))
    assert np.abs(hn[:n3] - c.hn).max() <= 4e-15 and np.abs(en[:n3] - c.en).max() <= 4e-15
    assert np.abs(r.view("pmldn")[:n3] - c.pmldn).max() <= 8e-15
    assert np.abs(r.view("pmlbn")[:n3] - c.pmlbn).max() <= 8e-15
    L.ref_set_user(0, C.cast(_usr(L, "userinc", "2ddielectric"), refrun.USERCB))
    c.step(40); r.step(40)
    num = np.sqrt(np.sum((c.hn - hn[:n3]) ** 2) + np.sum((c.en - en[:n3]) ** 2))
    den = np.sqrt(np.sum(c.hn ** 2) + np.sum(c.en ** 2))
    assert num / den <= 1e-12
    sol = [np.zeros(n) for _ in range(6)]
    tt = C.c_double(c.time)
    _usr(L, "usersol", "2ddielectric")(C.byref(tt), *[_dp(a) for a in sol])
    mh, me = u.usersol(c, c.time)
    for k in range(3):
        assert np.abs(sol[k] - c.comp(mh, k)).max() <= 4e-15
        assert np.abs(sol[3 + k] - c.comp(me, k)).max() <= 4e-15
    r.set("volvm1", c.volvm1)
    err = np.zeros(n)
    nn = C.c_int(n)
    for k in range(6):
        if c.tol["l2"][k] == 0:
            continue
        fld = np.ascontiguousarray(c.comp(c.hn if k < 3 else c.en, k % 3))
        l2, linf = C.c_double(), C.c_double()
        L.cem_error_(_dp(fld), _dp(sol[k]), _dp(err), C.byref(nn), C.byref(l2), C.byref(linf))
        assert l2.value <= c.tol["l2"][k] and linf.value <= c.tol["linf"][k], (k, l2.value, linf.value)
    r.close()


@pytest.mark.parametrize("imode", [1, 2])
def test_pin_2dboxpml_with_the_shipped_usr(imode):
    """tests/2dboxpml TE / TM driven by its own usrdat2 and usersrc (2D Gaussian source into hz
    or ez): bit-level PML state is covered by _assert_same on the oracle-side callback run, the
    .usr-driven run agrees to the round-off of libm vs numpy exp/sin"""
    from oracle import oracle as O
    c = cases.case_2dboxpml(imode, nx1=7, nel=(6, 6))
    r = refrun.ReferenceRun(c)
    raw = O.RefCase(c.mesh, c.nx1, imode=imode)
    r.put("xm1", raw.xm1); r.put("ym1", raw.ym1)
    _usr(r.L, "usrdat2", "2dboxpml")()
    assert np.array_equal(r.view("xm1")[:c.npts], c.xm1)
    assert np.array_equal(r.view("ym1")[:c.npts], c.ym1)
    r.L.ref_set_user(1, C.cast(_usr(r.L, "usersrc", "2dboxpml"), refrun.USERCB))
    c.step(30); r.step(30)
    n3 = 3 * c.npts
    num = np.sqrt(np.sum((c.hn - r.hn) ** 2) + np.sum((c.en - r.en) ** 2))
    den = np.sqrt(np.sum(c.hn ** 2) + np.sum(c.en ** 2))
    assert den > 1e-8 and num / den <= 1e-13
    for name in ("pmlbn", "pmldn"):
        a, b = getattr(c, name), r.view(name)[:n3]
        assert np.abs(a - b).max() <= 1e-13 * max(np.abs(a).max(), 1e-300), name
    r.close()
    # the same case with the oracle's own callback on both sides: bit for bit
    c = cases.case_2dboxpml(imode, nx1=7, nel=(6, 6))
    r = refrun.ReferenceRun(c)
    r.set_callback("usersrc", c.usersrc_fn)
    c.step(30); r.step(30)
    _assert_same(c, r)
    r.close()


# ---------------------------------------------------------------------------------------------
# tests/cylwave: unstructured mesh with circular-arc sides (curved, non-affine elements)
# ---------------------------------------------------------------------------------------------
def _put_corners(r, mesh, nelt):
    for name, arr in (("xc", mesh.xc), ("yc", mesh.yc), ("zc", mesh.zc)):
        v = r.view(name).reshape(-1)
        v[:] = 0.0
        v[:8 * nelt] = arr.reshape(-1)


def test_pin_cylwave_mesh_generation():
    """the whole chain from the reference's .rea corner coordinates to the GLL nodes of the
    curved elements, all reference code: usrdat of cylwave.usr (corners projected onto the
    radius found by geom_xyradius, src/cem_common.F:817-843), GENXYZ + ARCSRF
    (src/nek5_genxyz.F:2-108, 562-680) for the 80 sides flagged 'C', usrdat2 (z rescaled to
    2 pi xmax).  Bit for bit."""
    raw, _ = cases._load_mesh("cylwave_mesh.npz")       # corners exactly as in cylwave.rea
    c = cases.case_cylwave(nx1=8)
    mesh = c.mesh                                       # corners after the oracle's usrdat
    assert np.abs(mesh.xc - raw.xc).max() > 1e-7        # usrdat did move the outer corners
    r = refrun.ReferenceRun(c)
    _put_corners(r, raw, c.nelt)
    _usr(r.L, "usrdat", "cylwave")()
    assert np.array_equal(r.view("xc").reshape(-1)[:8 * c.nelt], mesh.xc.reshape(-1))
    assert np.array_equal(r.view("yc").reshape(-1)[:8 * c.nelt], mesh.yc.reshape(-1))
    r.L.geom_xyradius_.restype = C.c_double
    assert r.L.geom_xyradius_() == c.cyl_radius
    # CCURVE(12,lelt) character*1, CURVE(6,12,lelt)
    cc = r.view_char("ccurve")
    cc[:] = ord(" ")
    cv = r.view("curve")
    cv[:] = 0.0
    for e in range(c.nelt):
        for k in range(12):
            cc[k + 12 * e] = ord(mesh.ccurve[e][k])
            cv[6 * (k + 12 * e):6 * (k + 12 * e) + 5] = mesh.curve[e, k]
    n = c.nx1
    r.put("zgm1", np.concatenate([c.zgm1, c.zgm1, c.zgm1]))
    for k in ("ifgmsh3", "ifaxis"):
        try:
            r.set(k, 0)
        except KeyError:
            pass
    r.L.initds_()
    x, y, z = r.view("xm1"), r.view("ym1"), r.view("zm1")
    r.L.genxyz_(_dp(x), _dp(y), _dp(z), C.byref(C.c_int(n)), C.byref(C.c_int(n)),
                C.byref(C.c_int(n)))
    # nodes of the curved sides lie on the cylinder
    rad = np.sqrt(x[:c.npts] ** 2 + y[:c.npts] ** 2)
    assert abs(rad.max() - c.cyl_radius) < 1e-14 * c.cyl_radius
    _usr(r.L, "usrdat2", "cylwave")()
    assert np.array_equal(x[:c.npts], c.xm1) and np.array_equal(y[:c.npts], c.ym1)
    assert np.array_equal(z[:c.npts], c.zm1)
    r.close()


def test_pin_cylwave_time_stepping():
    """tests/cylwave at N=11 as shipped (50 curved elements, PEC wall, periodic in z, CFL 0.25):
    10 steps, fields and RK registers bit for bit; the analytic TM_01 mode stays within the
    userchk tolerances (5e-9 / 5e-8) evaluated by the reference's cem_error"""
    c, r = _pair(cases.case_cylwave())
    c.step(10); r.step(10)
    _assert_same(c, r)
    r.set("volvm1", c.volvm1)
    shn, sen = c.usersol(c, c.time)
    n = c.npts
    err = np.zeros(n)
    nn = C.c_int(n)
    for k in range(6):
        fld = np.ascontiguousarray(r.view("hn" if k < 3 else "en")[(k % 3) * n:(k % 3 + 1) * n])
        sol = np.ascontiguousarray(c.comp(shn if k < 3 else sen, k % 3))
        l2, linf = C.c_double(), C.c_double()
        r.L.cem_error_(_dp(fld), _dp(sol), _dp(err), C.byref(nn), C.byref(l2), C.byref(linf))
        assert l2.value <= 5e-9 and linf.value <= 5e-8, (k, l2.value, linf.value)
    r.close()


def test_pin_vtk_payload_against_the_references_writer_pieces():
    """Output hand-off: the oracle's VTK "VECTORS" payload equals what the reference assembles --
    its own vtk_nonswap_field (translated, src/io_dumpvtk.F:858-878) followed by the cast and
    byte swap of writefield4 / writefield4_double (src/io_co.c:443-456, 511-524) done with the
    reference's own swap_float_byte / swap_double_byte (src/io_util.c, compiled unchanged; the
    five-line loop around them needs MPI-IO and is restated here)."""
    c = cases.case_boxper((2, 2, 2), 4)
    c.step(2)
    r = refrun.ReferenceRun(c)
    L = r.L
    n = c.npts
    L.adjust_endian()
    for which in ("en", "hn"):
        f = getattr(c, which)
        out = np.zeros(3 * n)
        comps = [np.ascontiguousarray(f[k * n:(k + 1) * n]) for k in range(3)]
        L.vtk_nonswap_field_(_dp(comps[0]), _dp(comps[1]), _dp(comps[2]), _dp(out))
        # writefield4: (float) cast + swap_float_byte per value
        fl = out.astype(np.float32)
        fp = fl.ctypes.data_as(C.POINTER(C.c_float))
        for i in range(fl.size):
            L.swap_float_byte(C.byref(C.c_float.from_address(C.addressof(fp.contents) + 4 * i)))
        assert fl.tobytes() == c.vtk_payload(which, as_double=False)
        db = out.copy()
        dpp = db.ctypes.data_as(C.POINTER(C.c_double))
        for i in range(db.size):
            L.swap_double_byte(C.byref(C.c_double.from_address(C.addressof(dpp.contents) + 8 * i)))
        assert db.tobytes() == c.vtk_payload(which, as_double=True)
    assert np.abs(c.en).max() > 1e-3
    r.close()


# ---------------------------------------------------------------------------------------------
# the optional modal filter of cem_maxwell_op_rk (`if (iffilter) call q_filter(0.01)`, :342)
# ---------------------------------------------------------------------------------------------
@pytest.mark.parametrize("n", [4, 8, 9, 12, 16])
def test_pin_filter_matrix(n):
    """build_new_filter (src/nek5_filter.F:171-249, Gauss-Jordan inverse of the modal basis):
    the oracle's numpy restatement agrees to round-off; rows sum to 1 (constants pass)"""
    from oracle import oracle as O
    c = cases.case_boxper((3, 3, 3), n)
    r = refrun.ReferenceRun(c)
    intv = np.zeros(n * n)
    z = np.ascontiguousarray(c.zgm1)
    r.L.build_new_filter_(_dp(intv), _dp(z), C.byref(C.c_int(n)), C.byref(C.c_int(2)),
                          C.byref(C.c_double(0.01)), C.byref(C.c_int(1)))
    mine = O.build_new_filter(c.zgm1, 2, 0.01)
    assert np.abs(intv - mine).max() <= 2e-13
    assert np.abs(intv.reshape(n, n).T.sum(axis=1) - 1.0).max() <= 1e-13
    assert np.abs(intv.reshape(n, n).T - np.eye(n)).max() > 1e-4   # it does filter
    r.close()


@pytest.mark.parametrize("which", ["3d", "2d-te"])
def test_pin_filtered_time_stepping(which):
    """param(18) = 1: every time step ends with q_filter(0.01) = filterq on ex,ey,ez,hx,hy,hz
    (src/nek5_filter.F:68-74, 92-144).  Reference side: its own cem_maxwell_op_rk followed by its
    own filterq on the six components with the matrix from its own build_new_filter; the oracle
    uses the same matrix.  Fields and RK registers bit for bit after 6 filtered steps."""
    c = cases.case_boxper((3, 3, 3), 7, dt=-2e-3) if which == "3d" else cases.case_2dboxper(1, nx1=8)
    r = refrun.ReferenceRun(c)
    n, nz = c.nx1, (c.nx1 if c.ldim == 3 else 1)
    intv = np.zeros(n * n)
    z = np.ascontiguousarray(c.zgm1)
    r.L.build_new_filter_(_dp(intv), _dp(z), C.byref(C.c_int(n)), C.byref(C.c_int(2)),
                          C.byref(C.c_double(0.01)), C.byref(C.c_int(1)))
    c.filter = intv.copy()                         # the reference's own matrix on both sides
    w1 = np.zeros(c.nxyz * c.nelt); w2 = np.zeros(c.nxyz); ft = np.zeros(n * n)
    if3d, dmax = C.c_int(int(c.ldim == 3)), C.c_double()
    N = c.npts
    for _ in range(6):
        c.step(1)
        r.step(1)
        for name in ("en", "hn"):                  # q_filter's order: E first
            v = r.view(name)
            for k in range(3):
                r.L.filterq_(_dp(v[k * N:]), _dp(intv), C.byref(C.c_int(n)), C.byref(C.c_int(nz)),
                             _dp(w1), _dp(w2), _dp(ft), C.byref(if3d), C.byref(dmax))
    _assert_same(c, r)
    # and the filter is not a no-op on this run
    c0 = cases.case_boxper((3, 3, 3), 7, dt=-2e-3) if which == "3d" else cases.case_2dboxper(1, nx1=8)
    c0.step(6)
    assert np.abs(c0.en - c.en).max() > 1e-9
    r.close()


def test_pin_padded_size_layout():
    """A real SIZE file dimensions the arrays for lelt = lelg/lpmin + 3 > nelt elements
    (tests/3dboxper/SIZE:15), so hn(lpts1,3) etc. have a leading dimension larger than npts.
    The translated reference run with that layout (3 padding elements) equals the oracle run
    bit for bit -- this pins ReferenceRun's padded placement, which the drop-in test of the
    shim's leading-dimension handling (tests/test_gpu_zz_padded_size.py) relies on."""
    c = cases.case_3ddielectric(True, nx1=5, nel=(3, 6, 3))
    c.set_callback("userinc", lambda tt, *a: None)
    r = refrun.ReferenceRun(c, pad_elems=3)
    assert r.lpts1 == c.nxyz * (c.nelt + 3)
    c.step(5); r.step(5)
    for name in ("hn", "en", "khn", "ken", "pmlbn", "pmldn"):
        assert np.array_equal(r.field(name), getattr(c, name)), name
    assert np.abs(c.hn).max() > 1e-3
    r.close()


@pytest.mark.parametrize("imode", [1, 2])
def test_pin_central_flux_2d(imode):
    """param(19) = 1 (central flux, C0 = 0) through cem_maxwell_flux2d, TE and TM"""
    from oracle import oracle as O
    import math
    mesh = O.box_mesh((3, 3), ((0.0, 1.0),) * 2, ("P  ",) * 4)
    c = O.RefCase(mesh, 6, imode=imode, upwind=False,
                  usrdat2=lambda case: cases._rescale(case, (0.0, 0.0), (2 * math.pi, 2 * math.pi)))
    c.set_dt(-2e-3)
    shn, sen = cases.usersol_2dboxper(c, 0.0)
    c.hn[:] = shn; c.en[:] = sen
    r = refrun.ReferenceRun(c)
    assert r.get("ifcentral") == 1 and r.get("ifupwind") == 0
    c.step(5); r.step(5)
    _assert_same(c, r)
    r.close()


@pytest.mark.parametrize("kind", ["drude", "lorentz"])
def test_pin_ade_in_3d(kind):
    """cem_maxwell_drude / cem_maxwell_lorentz on a 3D mesh (the shipped tests use them in 2D):
    a periodic box whose lower half is dispersive, random initial currents; the reference's
    routine is called from a test-side usersrc on its own copies of the arrays"""
    from oracle import oracle as O
    c = cases.case_boxper((3, 4, 3), 5, dt=-2e-3)
    n = c.npts
    low = np.repeat((c.ym1 < np.pi).reshape(c.nelt, -1).all(axis=1), c.nxyz)
    index = np.nonzero(low)[0].astype(np.int32)
    npar, nj = (2, 3) if kind == "drude" else (3, 6)
    params = np.zeros(npar * n)
    params[0:n][low] = 0.3
    params[n:2 * n][low] = 4.0
    if npar == 3:
        params[2 * n:][low] = 2.5
    rng = np.random.default_rng(11)
    jn = np.zeros(nj * n)
    for k in range(nj):
        jn[k * n:(k + 1) * n][low] = 0.1 * rng.standard_normal(int(low.sum()))
    kjn, resjn = np.zeros(nj * n), np.zeros(nj * n)
    jr, kr, rr, pr = jn.copy(), kjn.copy(), resjn.copy(), params.copy()
    fn_o = c.L.ora_cem_maxwell_drude if kind == "drude" else c.L.ora_cem_maxwell_lorentz
    c.set_callback("usersrc", lambda tt, *res: fn_o(C.byref(c.s), O.dp(jn), O.dp(kjn), O.dp(resjn),
                                                      O.dp(params), O.ip(index), int(index.size)))
    r = refrun.ReferenceRun(c)
    fn_r = r.L.cem_maxwell_drude_ if kind == "drude" else r.L.cem_maxwell_lorentz_
    idx1 = (index + 1).astype(np.int32)
    nn = C.c_int(idx1.size)
    r.set_callback("usersrc", lambda tt, *res: fn_r(_dp(jr), _dp(kr), _dp(rr), _dp(pr),
                                                      idx1.ctypes.data_as(C.POINTER(C.c_int)),
                                                      C.byref(nn)))
    c.step(6); r.step(6)
    _assert_same(c, r)
    assert np.array_equal(jn, jr) and np.array_equal(kjn, kr) and np.abs(jn).max() > 1e-2
    r.close()


@pytest.mark.parametrize("imode", [1, 2])
@pytest.mark.parametrize("upwind", [True, False])
def test_pin_deformed_2d_mesh_with_pec_walls(imode, upwind):
    """a sheared, non-affine 2D mesh (all four metric terms, oblique face normals) with PEC walls,
    upwind and central flux, TE and TM: geometry from the oracle, time stepping bit for bit"""
    from oracle import oracle as O
    mesh = O.box_mesh((4, 4), ((-1.0, 1.0),) * 2, ("PEC",) * 4)

    def warp(case):
        x, y = case.xm1.copy(), case.ym1.copy()
        case.xm1[:] = x + 0.07 * np.sin(np.pi * y)
        case.ym1[:] = y + 0.05 * np.sin(np.pi * x) * np.cos(0.5 * np.pi * y)

    c = O.RefCase(mesh, 7, imode=imode, upwind=upwind, usrdat2=warp)
    c.set_dt(-2e-3)
    c.hn[:], c.en[:] = cases.usersol_2dboxpec(c, 0.0)
    r = refrun.ReferenceRun(c)
    c.step(6); r.step(6)
    _assert_same(c, r)
    r.close()


def test_pin_warped_3d_mesh_time_stepping():
    """sheared, non-affine 3D elements (all nine cofactors vary inside an element), periodic:
    10 steps bit for bit"""
    c = _geometry_case("warped")
    c.set_dt(-5e-4)
    # periodic in the unwarped coordinates: a smooth field of the reference box
    n = c.npts
    x, y, z = c.xm1, c.ym1, c.zm1
    c.hn[0:n] = np.sin(2 * np.pi * y) * np.cos(2 * np.pi * z)
    c.hn[n:2 * n] = np.cos(2 * np.pi * x)
    c.en[2 * n:] = np.sin(2 * np.pi * x) * np.sin(2 * np.pi * y)
    r = refrun.ReferenceRun(c)
    c.step(10); r.step(10)
    _assert_same(c, r)
    r.close()
