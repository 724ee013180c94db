"""Unit checks of the oracle's building blocks and of the product-side box generator against
the oracle's literal restatement of the reference setup."""
import numpy as np
import pytest

from helpers import arrays_from_refcase
from nekcem_b200.boxcase import BoxCase, dgll, gll, gllnid_box
from oracle import cases, oracle as O


@pytest.mark.parametrize("n", [2, 3, 5, 8, 9, 12, 16])
def test_gll_quadrature(n):
    z, w = O.zwgll(n)
    assert z[0] == -1.0 and z[-1] == 1.0 and np.all(np.diff(z) > 0)
    assert abs(w.sum() - 2.0) < 1e-13
    for p in range(0, 2 * n - 2):  # exact for degree <= 2n-3
        exact = 0.0 if p % 2 else 2.0 / (p + 1)
        assert abs(np.dot(w, z ** p) - exact) < 1e-12
    z2, w2 = gll(n)
    assert np.max(np.abs(z - z2)) < 1e-14 and np.max(np.abs(w - w2)) < 1e-14


@pytest.mark.parametrize("n", [3, 8, 9, 16])
def test_dgll(n):
    z, _ = O.zwgll(n)
    d, dt = O.dgll(z)
    D = d.reshape(n, n).T  # column-major -> D[i, j]
    assert np.allclose(dt.reshape(n, n), D, atol=0)  # dt(j,i) = d(i,j)
    for p in range(n):
        exact = p * z ** (p - 1) if p > 0 else 0 * z
        assert np.max(np.abs(D @ z ** p - exact)) < 1e-10
    assert np.max(np.abs(D - dgll(z))) < 1e-12


def test_mxm_left_to_right():
    rng = np.random.default_rng(0)
    a = rng.standard_normal((5, 7)); b = rng.standard_normal((7, 4))
    c = np.zeros(20)
    O.lib().ora_mxm(O.dp(np.asfortranarray(a).reshape(-1, order="F").copy()), 5,
                    O.dp(np.asfortranarray(b).reshape(-1, order="F").copy()), 7, O.dp(c), 4)
    ref = np.zeros((5, 4))
    for i in range(5):
        for j in range(4):
            s = a[i, 0] * b[0, j]
            for k in range(1, 7):
                s = s + a[i, k] * b[k, j]
            ref[i, j] = s
    assert np.array_equal(c.reshape(4, 5).T, ref)  # bit-exact summation order


def test_gs_pairwise_sum():
    ids = np.array([5, 0, 7, 5, 9, 7, 0], dtype=np.int64)
    L = O.lib()
    import ctypes as C
    g = L.ora_gs_setup(ids.ctypes.data_as(C.POINTER(C.c_longlong)), ids.size)
    u = np.arange(14, dtype=np.float64)  # two fields, stride 7
    L.ora_gs_op_fields(g, O.dp(u), 7, 2, 1)
    assert list(u[:7]) == [3, 1, 7, 3, 4, 7, 6]
    assert list(u[7:]) == [17, 8, 21, 17, 11, 21, 13]
    L.ora_gs_free(g)


def test_cemface_and_face_ids_pair_coincident_points():
    c = cases.case_3dboxper()
    # paired face points are geometrically identical modulo the 2*pi period
    order = np.argsort(c.glo_num, kind="stable")
    g = c.glo_num[order]
    assert np.all(g[0::2] == g[1::2]) and np.all(g[0:-2:2] != g[2::2])
    a, b = c.cemface[order[0::2]], c.cemface[order[1::2]]
    for x in (c.xm1, c.ym1, c.zm1):
        d = np.abs(x[a] - x[b])
        d = np.minimum(d, np.abs(d - 2 * np.pi))
        assert d.max() < 1e-12
    # unit outward normals of a pair are opposite, areas equal
    ja, jb = order[0::2], order[1::2]
    assert np.max(np.abs(c.unxm[ja] + c.unxm[jb])) < 1e-12
    assert np.max(np.abs(c.aream[ja] - c.aream[jb])) < 1e-13


def test_boxcase_matches_oracle_setup():
    c = cases.case_boxper((3, 4, 5), 6)
    b = BoxCase((3, 4, 5), 6)
    A, B = arrays_from_refcase(c), b.arrays()
    for k, a in A.items():
        if k in ("glo_num", "cempec", "pmlptr", "volvm1") or k not in B:
            continue
        a = np.asarray(a); bb = np.asarray(B[k])
        assert a.shape == bb.shape, k
        assert np.max(np.abs(a - bb)) <= 1e-12 * max(1.0, np.max(np.abs(a))), k
    assert abs(A["volvm1"] - B["volvm1"]) < 1e-10


def test_pencil_map():
    g = gllnid_box(4, 4, 8, 2)
    assert np.array_equal(np.bincount(g), [64, 64])
    assert np.all(g[: 4 * 4 * 4] == 0) and np.all(g[4 * 4 * 4:] == 1)  # z-slabs
    g = gllnid_box(2, 3, 5, 4)  # uneven: 15 pencils over 4 ranks, boustrophedon order
    assert sorted(np.bincount(g) // 2) == [3, 4, 4, 4]


def test_rotated_element_frames_leave_the_solution_unchanged():
    """Relabelling every element's local frame by a proper rotation must not change the physical
    solution: checks the oracle's orientation handling in face_glo_num (setup_dgds2 semantics)
    and in the metric/normal computation independently of any other implementation."""
    from oracle import cases
    nel, nx1 = (3, 3, 3), 5
    c0 = cases.case_boxper(nel, nx1, dt=-1e-3)
    c1, rots = cases.case_boxper_rotated(nel, nx1, dt=-1e-3)
    assert len(set(rots)) == 24
    m = cases.rotated_node_map(nx1, c0.nelt, rots)
    assert np.abs(c1.xm1[m] - c0.xm1).max() < 1e-14
    c0.step(10); c1.step(10)
    n = c0.npts
    for a, b in ((c1.hn, c0.hn), (c1.en, c0.en)):
        assert np.abs(a.reshape(3, n)[:, m] - b.reshape(3, n)).max() < 1e-13


@pytest.mark.parametrize("which", ["3ddielectric", "2ddielectric-te", "2ddielectric-tm", "drude",
                                   "lorentz", "3dgraphene", "2dgraphene-te", "2dgraphene-tm"])
def test_planewave_parameters_reproduce_usersol(which):
    """tests/helpers.py: planewave_args (the arguments of the device-side plane-wave usersol,
    nekcem_b200_error_sums_planewave) evaluated with numpy reproduces the usersol of every
    layered-media case -- i.e. the two-half-space parameterisation is exact, PML decay, complex
    wavenumber of the Drude metal and complex graphene coefficients included."""
    import sys, os
    sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
    from helpers import planewave_args, planewave_numpy
    from oracle import cases
    mk = {"3ddielectric": lambda: cases.case_3ddielectric(True),
          "2ddielectric-te": lambda: cases.case_2ddielectric(1, True),
          "2ddielectric-tm": lambda: cases.case_2ddielectric(2, False),
          "drude": cases.case_drude, "lorentz": cases.case_lorentz,
          "3dgraphene": lambda: cases.case_3dgraphene(nel=(3, 12, 3)),
          "2dgraphene-te": lambda: cases.case_2dgraphene(1),
          "2dgraphene-tm": lambda: cases.case_2dgraphene(2)}[which]
    c = mk()
    a = planewave_args(c)
    for tt in (0.0, 0.37, 2.5):
        sh, se = c.usersol(c, tt)
        ph, pe = planewave_numpy(c, a, tt)
        assert np.abs(sh).max() > 0.5
        assert np.abs(sh - ph).max() <= 1e-15 and np.abs(se - pe).max() <= 1e-15


def test_boxcase_two_material_arrays_match_the_oracles_setup():
    """nekcem_b200/boxcase.py generates the benchmark meshes analytically; its two-material variant
    (masses, impedances) must equal what the oracle's restatement of cem_maxwell_materials and the
    inverse-mass setup (src/cem_maxwell.F:183-186, 262-325) builds for the same box"""
    import numpy as np
    from nekcem_b200.boxcase import BoxCase
    from oracle import oracle as O
    nel, nx1 = (3, 4, 3), 5
    mesh = O.box_mesh(nel, ((0.0, 2 * np.pi),) * 3, ("P  ",) * 6)

    def uservp(case):
        e = np.arange(case.nelt)
        ey = (e // nel[0]) % nel[1]
        case.permittivity[:] = np.repeat(np.where(ey >= nel[1] // 2, 4.0, 1.0), case.nxyz)
        case.permeability[:] = 1.0

    c = O.RefCase(mesh, nx1, upwind=True, uservp=uservp)
    b = BoxCase(nel, nx1, eps_upper=4.0)
    for k in ("Y_0", "Y_1", "Z_0", "Z_1"):
        assert np.array_equal(getattr(c, k), b.array(k)), k
    for k in ("hbm1", "ebm1", "bmn", "permittivity", "permeability"):
        assert np.allclose(getattr(c, k), b.array(k), rtol=1e-12, atol=0), k
    assert BoxCase(nel, nx1, pml="layers").array("pmlptr").size == 2 * 2 * 3 * 3
    assert BoxCase(nel, nx1, pml=True).array("pmlsigma").size == 3 * b.npts


def test_boxcase2d_arrays_match_the_oracles_setup():
    """the analytic 2D periodic box of nekcem_b200/boxcase.py (benchmark meshes for the TE / TM
    path) against the oracle's setup of tests/2dboxper on the same 3 x 4 box"""
    import numpy as np
    from nekcem_b200.boxcase import BoxCase2D
    from oracle import cases
    for imode in (1, 2):
        c = cases.case_2dboxper(imode, nx1=5, nel=(3, 4))
        b = BoxCase2D((3, 4), 5, imode=imode)
        for k in ("dxm1", "w3mn", "rxmn", "rymn", "sxmn", "symn", "tzmn", "bmn", "unxm", "unym",
                  "aream", "Y_0", "Y_1", "Z_0", "Z_1", "hn", "en"):
            assert np.allclose(getattr(c, k), b.array(k), rtol=0, atol=1e-13), (imode, k)
        assert np.allclose(c.hbm1, b.array("hbm1"), rtol=1e-12)

        def pairs(g):
            d = {}
            for i, v in enumerate(g):
                d.setdefault(int(v), []).append(i)
            return sorted(tuple(v) for v in d.values())
        assert pairs(c.glo_num) == pairs(b.array("glo_num"))


def test_sheet_rk_registers_are_ill_conditioned():
    """Why the GPU tests compare the RK registers kfjn of the graphene sheet ODEs with 2e-11 instead
    of 1e-12: flip the last bit of the oracle's OWN initial fields (tests/2dgraphene as shipped, TE,
    200 steps) and its fields move by ~1e-15, its sheet currents by ~2e-15, but its kfjn by ~3e-12
    -- the critical-point residual -a21*j - a22*j' + b2*f with a21 ~ 4.6e5 is a difference of terms
    1e3..1e4 times its size.  No implementation that differs by round-off can agree better."""
    import numpy as np
    from oracle import cases

    def run(perturb):
        c = cases.case_2dgraphene(1)
        if perturb:
            rng = np.random.default_rng(1)
            for a in (c.hn, c.en):
                a *= 1.0 + rng.integers(-1, 2, a.size) * 2.2e-16
        c.step(200)
        u = c.user
        m = np.zeros(c.nxzfl, bool)
        m[u.graphindex] = True
        m18 = np.tile(m, 18)
        return np.concatenate([c.hn, c.en]), u.fjn[m18].copy(), u.kfjn[m18].copy()

    rel = lambda a, b: float(np.sqrt(np.sum((a - b) ** 2) / np.sum(b ** 2)))
    f0, j0, k0 = run(False)
    f1, j1, k1 = run(True)
    assert rel(f1, f0) <= 1e-14 and rel(j1, j0) <= 1e-13
    assert 5e-13 <= rel(k1, k0) <= 2e-11
