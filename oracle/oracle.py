"""CPU ORACLE driver (test infrastructure, NOT product code).

ctypes binding of ``nekcem_oracle.c`` plus numpy restatements of the reference's
setup-time routines (mesh, face numbering, BC flags, materials, PML layout) that feed
the hot path.  Each function cites the NekCEM file:line it follows.

PARITY STATUS: PINNED against the reference's own code executed in this container.
`oracle/_ref/libnekcem_ref.so` (recipe oracle/build_ref.py) holds the reference's Fortran
routines of the path -- translated statement by statement from /root/reference/src by
oracle/f2c_lite.py, since no Fortran compiler exists here -- linked with the reference's own
src/jl gather-scatter library compiled unchanged.  tests/test_reference_pin.py requires this
oracle to reproduce it BIT FOR BIT (fields, RK registers, PML/ADE state) on every shipped case
of the path and for nx1 = 2..17, and pins the setup it feeds the path (GLL nodes/weights/D,
cofactors, Jacobian, mass, face areas/normals, cemface, impedances, PEC list, PML tags/extents/
sigma, error norms, dxmin) and the .usr analytic solutions the same way.  What stays unpinned:
the reference compiled by a real Fortran compiler (gfortran/ifort code generation, e.g. FMA
contraction with -march=native) -- a <= 1e-15 relative effect per operation.  The reference's
known-answer tolerances are checked besides (tests/test_oracle_kat.py).

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs
may import this module.  All index arrays are 0-based.
"""
from __future__ import annotations

import ctypes as C
import math
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIBPATH = os.path.join(_HERE, "_build", "libnekcem_oracle.so")
_lib = None

c_dp = C.POINTER(C.c_double)
c_ip = C.POINTER(C.c_int)
USERCB = C.CFUNCTYPE(None, C.c_double, c_dp, c_dp, c_dp, c_dp, c_dp, c_dp, C.c_void_p)


def build(force: bool = False) -> str:
    """Compile the C restatement with the committed recipe (oracle/Makefile)."""
    src = os.path.join(_HERE, "nekcem_oracle.c")
    if force or not os.path.exists(_LIBPATH) or os.path.getmtime(_LIBPATH) < os.path.getmtime(src):
        subprocess.check_call(["make", "-C", _HERE], stdout=subprocess.DEVNULL)
    return _LIBPATH


class OraState(C.Structure):
    _fields_ = [
        ("ldim", C.c_int), ("nx1", C.c_int), ("nelt", C.c_int), ("imode", C.c_int),
        ("nxyz", C.c_int), ("nxzf", C.c_int), ("nfaces", C.c_int), ("npts", C.c_int),
        ("nxzfl", C.c_int),
        ("ifupwind", C.c_int), ("ifcentral", C.c_int), ("ifpml", C.c_int), ("ifpec", C.c_int),
        ("dt", C.c_double), ("time", C.c_double), ("rktime", C.c_double),
        ("rkstep", C.c_int), ("istep", C.c_int),
        ("rk4a", C.c_double * 5), ("rk4b", C.c_double * 5), ("rk4c", C.c_double * 6),
        ("dxm1", c_dp), ("dxtm1", c_dp), ("w3mn", c_dp),
        ("rxmn", c_dp), ("rymn", c_dp), ("rzmn", c_dp), ("sxmn", c_dp), ("symn", c_dp),
        ("szmn", c_dp), ("txmn", c_dp), ("tymn", c_dp), ("tzmn", c_dp), ("bmn", c_dp),
        ("unxm", c_dp), ("unym", c_dp), ("unzm", c_dp), ("aream", c_dp),
        ("cemface", c_ip), ("ncemface", C.c_int),
        ("cempec", c_ip), ("ncempec", C.c_int),
        ("gsh_face", C.c_void_p),
        ("hn", c_dp), ("en", c_dp), ("khn", c_dp), ("ken", c_dp), ("reshn", c_dp),
        ("resen", c_dp),
        ("fhn", c_dp), ("fen", c_dp), ("srflx", c_dp),
        ("hbm1", c_dp), ("ebm1", c_dp),
        ("Y_0", c_dp), ("Y_1", c_dp), ("Z_0", c_dp), ("Z_1", c_dp),
        ("permittivity", c_dp), ("permeability", c_dp),
        ("maxpml", C.c_int), ("pmlptr", c_ip),
        ("pmlsigma", c_dp), ("pmlbn", c_dp), ("pmldn", c_dp), ("respmlbn", c_dp),
        ("respmldn", c_dp), ("respmlhn", c_dp), ("respmlen", c_dp), ("kpmlbn", c_dp),
        ("kpmldn", c_dp),
        ("userinc", USERCB), ("usersrc", USERCB), ("userfsrc", USERCB),
        ("ctx", C.c_void_p),
    ]


def lib():
    global _lib
    if _lib is None:
        build()
        L = C.CDLL(_LIBPATH)
        L.ora_zwgll.argtypes = [c_dp, c_dp, C.c_int]
        L.ora_dgll.argtypes = [c_dp, c_dp, c_dp, C.c_int]
        L.ora_mxm.argtypes = [c_dp, C.c_int, c_dp, C.c_int, c_dp, C.c_int]
        L.ora_genxyz.argtypes = [C.c_int, C.c_int, C.c_int] + [c_dp] * 7
        L.ora_geom.argtypes = [C.c_int, C.c_int, C.c_int] + [c_dp] * 22
        L.ora_set_fc_ptr.argtypes = [C.c_int, C.c_int, C.c_int, c_ip]
        L.ora_gs_setup.argtypes = [C.POINTER(C.c_longlong), C.c_int]
        L.ora_gs_setup.restype = C.c_void_p
        L.ora_gs_free.argtypes = [C.c_void_p]
        L.ora_gs_op_fields.argtypes = [C.c_void_p, c_dp, C.c_int, C.c_int, C.c_int]
        sp = C.POINTER(OraState)
        for name in ("ora_cem_maxwell", "ora_restrict_to_face", "ora_flux",
                     "ora_add_flux_to_res", "ora_pml_step", "ora_invqmass",
                     "ora_cem_maxwell_op", "ora_rk_storage", "ora_cem_maxwell_op_rk"):
            getattr(L, name).argtypes = [sp]
            getattr(L, name).restype = None
        L.ora_rk_maxwell_ab.argtypes = [sp, C.c_int]
        L.ora_advance.argtypes = [sp, C.c_int]
        L.ora_rk4_upd.argtypes = [c_dp, c_dp, c_dp, C.c_double, C.c_double, C.c_double, C.c_int]
        L.ora_cem_maxwell_drude.argtypes = [sp, c_dp, c_dp, c_dp, c_dp, c_ip, C.c_int]
        L.ora_cem_maxwell_lorentz.argtypes = [sp, c_dp, c_dp, c_dp, c_dp, c_ip, C.c_int]
        L.ora_q_filter.argtypes = [sp, c_dp]
        L.ora_cem_graphene_current.argtypes = [sp, c_dp, c_dp, c_dp, c_dp, c_dp, c_ip, C.c_int]
        L.ora_cem_error.argtypes = [c_dp, c_dp, c_dp, C.c_int, c_dp, C.c_double, c_dp, c_dp]
        L.ora_get_dxmin.argtypes = [C.c_int, C.c_int, C.c_int, c_dp, c_dp, c_dp]
        L.ora_get_dxmin.restype = C.c_double
        L.ora_state_size.restype = C.c_int
        L.ora_num_threads.restype = C.c_int
        assert L.ora_state_size() == C.sizeof(OraState), "ora_state layout mismatch"
        _lib = L
    return _lib


def dp(a: np.ndarray):
    assert a.dtype == np.float64 and a.flags["C_CONTIGUOUS"]
    return a.ctypes.data_as(c_dp)


def ip(a: np.ndarray):
    assert a.dtype == np.int32 and a.flags["C_CONTIGUOUS"]
    return a.ctypes.data_as(c_ip)


# ----------------------------------------------------------------------------------
# speclib
# ----------------------------------------------------------------------------------
def zwgll(n: int):
    """GLL nodes/weights, src/nek5_speclib.F:107-122."""
    z = np.zeros(n)
    w = np.zeros(n)
    lib().ora_zwgll(dp(z), dp(w), n)
    return z, w


def dgll(z: np.ndarray):
    """Derivative matrix D and transpose (column-major flat), src/nek5_speclib.F:807-840."""
    n = z.size
    d = np.zeros(n * n)
    dt = np.zeros(n * n)
    lib().ora_dgll(dp(d), dp(dt), dp(z), n)
    return d, dt


# ----------------------------------------------------------------------------------
# meshes
# ----------------------------------------------------------------------------------
EFACE = (4, 2, 1, 3, 5, 6)  # Ed's (symmetric) face -> preprocessor slot, nek5_connect11.F:1067-1072
EFACE1 = (3, 2, 4, 1, 5, 6)  # inverse, :1076-1081


class Mesh:
    """Element corner coordinates (preprocessor corner order), BC strings per face slot
    (preprocessor order 1..6 = -y,+x,+y,-x,-z,+z), global vertex ids (symmetric order
    l = i + 2j + 4k) and, for box meshes, the box shape."""

    def __init__(self, ldim, xc, yc, zc, cbc, vertex, nelbox=None, ccurve=None, curve=None):
        # curved sides (src/INPUT: CCURVE(12,lelt), CURVE(6,12,lelt)): ccurve[e][edge] is the
        # 1-char flag (' ' straight, 'C' circular arc of radius curve[e,edge,0]); None = none
        self.ccurve, self.curve = ccurve, curve
        self.ldim = ldim
        self.xc, self.yc, self.zc = xc, yc, zc
        self.cbc = cbc  # list[nelt][2*ldim] of 3-char strings
        self.vertex = vertex  # (nelt, 2**ldim) int64
        self.nelt = xc.shape[0]
        self.nelbox = nelbox


def _distribution(nel, x0, x1, gain=1.0):
    """get_xyz_distribution, src/nek5_genbox.F:839-869 (negative counts = uniform/geometric)."""
    x = np.zeros(nel + 1)
    if gain == 1.0:
        dx = (x1 - x0) / nel
        for i in range(1, nel + 1):
            x[i] = x0 + i * dx
        x[0] = x0
        x[nel] = x1
    else:
        dx = (x1 - x0) / nel
        x[0] = 0.0
        for i in range(1, nel + 1):
            x[i] = x[i - 1] + dx
            dx = gain * dx
        scale = (x1 - x0) / x[nel]
        x = x * scale + x0
        x[nel] = x1
    return x


def box_mesh(nel, lo_hi, bcs, gain=(1.0, 1.0, 1.0)) -> Mesh:
    """.box mesh: makebox (src/nek5_genbox.F:402-560) + gen_gtp_vertex (:83-153).

    nel = (nelx, nely[, nelz]); lo_hi = ((x0,x1),(y0,y1)[,(z0,z1)]);
    bcs in .box order (-x,+x,-y,+y[,-z,+z]), e.g. ('P  ',)*6.
    Elements are ordered x fastest, then y, then z.
    """
    ldim = len(nel)
    nelx, nely = nel[0], nel[1]
    nelz = nel[2] if ldim == 3 else 1
    xs = _distribution(nelx, *lo_hi[0], gain[0])
    ys = _distribution(nely, *lo_hi[1], gain[1])
    zs = _distribution(nelz, *lo_hi[2], gain[2]) if ldim == 3 else np.zeros(2)
    nelt = nelx * nely * nelz
    nc = 2 ** ldim
    xc = np.zeros((nelt, nc)); yc = np.zeros((nelt, nc)); zc = np.zeros((nelt, nc))
    cbc = []
    vertex = np.zeros((nelt, nc), dtype=np.int64)

    def jjnt(n):
        return list(range(1, n + 1))

    nptsx, nptsy, nptsz = nelx + 1, nely + 1, nelz + 1
    indx, indy, indz = jjnt(nptsx), jjnt(nptsy), jjnt(nptsz)
    if bcs[0] == "P  ":
        indx[nptsx - 1] = 1
        nptsx = nelx
    if bcs[2] == "P  ":
        indy[nptsy - 1] = 1
        nptsy = nely
    if ldim == 3 and bcs[4] == "P  ":
        indz[nptsz - 1] = 1
        nptsz = nelz

    e = 0
    for ez in range(1, nelz + 1):
        for ey in range(1, nely + 1):
            for ex in range(1, nelx + 1):
                xc[e, 0:4] = (xs[ex - 1], xs[ex], xs[ex], xs[ex - 1])
                yc[e, 0:4] = (ys[ey - 1], ys[ey - 1], ys[ey], ys[ey])
                if ldim == 3:
                    xc[e, 4:8] = xc[e, 0:4]
                    yc[e, 4:8] = yc[e, 0:4]
                    zc[e, 0:4] = zs[ez - 1]
                    zc[e, 4:8] = zs[ez]
                cbc1 = bcs[0] if ex == 1 else "E  "
                cbc2 = bcs[1] if ex == nelx else "E  "
                cbc3 = bcs[2] if ey == 1 else "E  "
                cbc4 = bcs[3] if ey == nely else "E  "
                row = [cbc3, cbc2, cbc4, cbc1]
                if ldim == 3:
                    cbc5 = bcs[4] if ez == 1 else "E  "
                    cbc6 = bcs[5] if ez == nelz else "E  "
                    row += [cbc5, cbc6]
                cbc.append(row)
                l = 0
                for k in range(2 if ldim == 3 else 1):
                    for j in range(2):
                        for i in range(2):
                            v = indx[i + ex - 1] + (indy[j + ey - 1] - 1) * nptsx
                            if ldim == 3:
                                v += (indz[k + ez - 1] - 1) * nptsx * nptsy
                            vertex[e, l] = v
                            l += 1
                e += 1
    return Mesh(ldim, xc, yc, zc, cbc, vertex, nelbox=(nelx, nely, nelz))


def mesh_from_arrays(ldim, xc, yc, zc, cbc, vertex, ccurve=None, curve=None) -> Mesh:
    return Mesh(ldim, np.ascontiguousarray(xc, dtype=np.float64),
                np.ascontiguousarray(yc, dtype=np.float64),
                np.ascontiguousarray(zc, dtype=np.float64),
                [list(r) for r in cbc], np.ascontiguousarray(vertex, dtype=np.int64),
                ccurve=None if ccurve is None else [[str(c) for c in r] for r in ccurve],
                curve=None if curve is None else np.ascontiguousarray(curve, dtype=np.float64))


# ----------------------------------------------------------------------------------
# face-point global ids (what the reference hands to gs_setup(gsh_face,...))
# ----------------------------------------------------------------------------------
def _face_corner_vertices(ldim):
    """For each preprocessor face slot, the symmetric-vertex indices of the face corners on
    the face lattice (a,b): list of ((a,b)->vertex index l=i+2j+4k)."""
    out = []
    if ldim == 3:
        fixed = [("j", 0), ("i", 1), ("j", 1), ("i", 0), ("k", 0), ("k", 1)]
    else:
        fixed = [("j", 0), ("i", 1), ("j", 1), ("i", 0)]
    for ax, val in fixed:
        corners = {}
        nb = 2 if ldim == 3 else 1
        for b in range(nb):
            for a in range(2):
                if ldim == 3:
                    if ax == "i":
                        i, j, k = val, a, b
                    elif ax == "j":
                        i, j, k = a, val, b
                    else:
                        i, j, k = a, b, val
                else:
                    if ax == "i":
                        i, j, k = val, a, 0
                    else:
                        i, j, k = a, val, 0
                corners[(a, b)] = i + 2 * j + 4 * k
        out.append(corners)
    return out


def face_glo_num(mesh: Mesh, nx1: int) -> np.ndarray:
    """Global ids of the face points, semantics of setup_dgds2 / set_vert2 /
    iface_vert_int8 (src/nek5_connect11.F:2179-2347, 453-600): every face point carries an
    id shared by exactly the coincident point on the neighbouring element's face; the
    orientation is resolved from the global vertex ids ("count from the minimum global
    vertex").  Unpaired (physical boundary) faces get ids nobody else has.
    Layout: (nelt, nfaces, nxzf) flattened, slot order = preprocessor faces."""
    ldim = mesh.ldim
    n = nx1
    nfaces = 2 * ldim
    nxzf = n * (n if ldim == 3 else 1)
    corners = _face_corner_vertices(ldim)
    glo = np.zeros((mesh.nelt, nfaces, nxzf), dtype=np.int64)
    face_ids = {}
    a_idx = np.arange(n)
    for e in range(mesh.nelt):
        v = mesh.vertex[e]
        for f in range(nfaces):
            cv = {ab: int(v[l]) for ab, l in corners[f].items()}
            key = tuple(sorted(cv.values()))
            fid = face_ids.setdefault(key, len(face_ids))
            if ldim == 3:
                # canonical frame: origin at the corner with the smallest vertex id; first
                # axis towards the adjacent corner with the smaller id.
                (a0, b0) = min(cv, key=lambda ab: cv[ab])
                va = cv[(1 - a0, b0)]  # neighbour along a
                vb = cv[(a0, 1 - b0)]  # neighbour along b
                A, B = np.meshgrid(a_idx, a_idx, indexing="ij")  # A[a,b]=a
                pa = A if a0 == 0 else (n - 1 - A)
                pb = B if b0 == 0 else (n - 1 - B)
                if va < vb:
                    canon = pa + n * pb
                else:
                    canon = pb + n * pa
                # face point index p = a + n*b
                ids = np.zeros(nxzf, dtype=np.int64)
                ids[(A + n * B).ravel()] = (fid * nxzf + canon + 1).ravel()
            else:
                a0 = 0 if cv[(0, 0)] < cv[(1, 0)] else 1
                pa = a_idx if a0 == 0 else (n - 1 - a_idx)
                ids = fid * nxzf + pa + 1
            glo[e, f, :] = ids
    return glo.reshape(-1)


# ----------------------------------------------------------------------------------
# The reference case: COMMON-block arrays + step driver
def build_new_filter(zpts: np.ndarray, kut: int, wght: float) -> np.ndarray:
    """build_new_filter src/nek5_filter.F:171-249: the 1D modal filter F = V D V^-1 in the basis
    phi_1 = L_0, phi_2 = L_1, phi_k = L_{k-1} - L_{k-3}, with transfer function d_k = 1 below
    k0 = nx - kut and 1 - wght (k-k0)^2 / kut^2 above (quadratic roll-off of the top kut modes).
    Column-major n x n, flat.  numpy's solve in place of the reference's Gauss-Jordan inverse
    (setup; agreement ~1e-14, tests/test_reference_pin.py) -- in the drop-in the matrix comes
    from the host's own build_new_filter."""
    nx = zpts.size
    n = nx - 1
    pht = np.zeros((nx, nx))  # pht[k, j] = phi_k(z_j)
    for j, z in enumerate(zpts):
        L = np.zeros(nx + 1)
        L[0] = 1.0
        L[1] = z
        for k in range(2, n + 1):  # legendre_poly src/nek5_grad.F:272-289
            L[k] = ((2 * k - 1) * z * L[k - 1] - (k - 1) * L[k - 2]) / k
        pht[0, j] = L[0]
        pht[1, j] = L[1]
        for k in range(2, nx):
            pht[k, j] = L[k] - L[k - 2]
    phi = pht.T.copy()        # phi[j, k] = phi_k(z_j): V
    diag = np.eye(nx)
    k0 = nx - kut
    for k in range(k0 + 1, nx + 1):
        amp = wght * (k - k0) * (k - k0) / (kut * kut)
        diag[k - 1, k - 1] = 1.0 - amp
    intv = phi @ (diag @ np.linalg.inv(phi))
    return np.ascontiguousarray(intv.T).reshape(-1)  # column-major flat


# ----------------------------------------------------------------------------------
class RefCase:
    """Restates cem_init/cem_solve setup for the Maxwell RK path (src/cem_drive.F:17-190,
    src/cem_maxwell.F:65-191) on a given mesh, and drives the C step."""

    def __init__(self, mesh: Mesh, nx1: int, imode: int | None = None, upwind: bool = True,
                 usrdat2=None, uservp=None, param: dict | None = None, omp_threads=None):
        L = lib()
        self.L = L
        self.mesh = mesh
        self.param = dict(param or {})
        ldim = mesh.ldim
        self.ldim, self.nx1, self.nelt = ldim, nx1, mesh.nelt
        nz1 = nx1 if ldim == 3 else 1
        self.nxyz = nx1 * nx1 * nz1
        self.nxzf = nx1 * nz1
        self.nfaces = 2 * ldim
        self.npts = self.nxyz * self.nelt
        self.nxzfl = self.nxzf * self.nfaces * self.nelt
        # imode: param(4) (1 TE, 2 TM) in 2D, 3 in 3D (src/cem_param.F:48-92)
        self.imode = 3 if ldim == 3 else (imode or 1)
        npts, nxzfl, nelt = self.npts, self.nxzfl, self.nelt

        # genwz (src/nek5_coef.F:225-262)
        self.zgm1, self.wxm1 = zwgll(nx1)
        self.dxm1, self.dxtm1 = dgll(self.zgm1)

        # gengeom: genxyz
        self.xm1 = np.zeros(npts); self.ym1 = np.zeros(npts); self.zm1 = np.zeros(npts)
        L.ora_genxyz(ldim, nx1, nelt, dp(self.zgm1), dp(np.ascontiguousarray(mesh.xc)),
                     dp(np.ascontiguousarray(mesh.yc)), dp(np.ascontiguousarray(mesh.zc)),
                     dp(self.xm1), dp(self.ym1), dp(self.zm1))
        # curved sides: the ARCSRF loop of GENXYZ (src/nek5_genxyz.F:669-674)
        if mesh.ccurve is not None:
            self._arcsrf()
        # usrdat2 (user rescale) then geom_reset
        if usrdat2 is not None:
            usrdat2(self)
        z = lambda n: np.zeros(n)
        self.rxmn, self.rymn, self.rzmn = z(npts), z(npts), z(npts)
        self.sxmn, self.symn, self.szmn = z(npts), z(npts), z(npts)
        self.txmn, self.tymn, self.tzmn = z(npts), z(npts), z(npts)
        self.jacm, self.bmn, self.w3mn = z(npts), z(npts), z(self.nxyz)
        self.aream, self.unxm, self.unym, self.unzm = z(nxzfl), z(nxzfl), z(nxzfl), z(nxzfl)
        L.ora_geom(ldim, nx1, nelt, dp(self.dxm1), dp(self.dxtm1), dp(self.wxm1),
                   dp(self.xm1), dp(self.ym1), dp(self.zm1),
                   dp(self.rxmn), dp(self.rymn), dp(self.rzmn), dp(self.sxmn), dp(self.symn),
                   dp(self.szmn), dp(self.txmn), dp(self.tymn), dp(self.tzmn), dp(self.jacm),
                   dp(self.bmn), dp(self.w3mn), dp(self.aream), dp(self.unxm), dp(self.unym),
                   dp(self.unzm))
        self.volvm1 = float(np.sum(self.bmn))  # src/nek5_coef.F:973

        # setup_topo -> setup_dgds2 -> gs_setup(gsh_face)
        self.glo_num = face_glo_num(mesh, nx1)
        self.gsh = L.ora_gs_setup(self.glo_num.ctypes.data_as(C.POINTER(C.c_longlong)), nxzfl)

        # setlog (src/nek5_bdry.F:68-92)
        flat = [cb for row in mesh.cbc for cb in row]
        self.ifpec = any(cb in ("PEC", "pec") for cb in flat)
        self.ifpml = any(cb in ("PML", "pml") for cb in flat)

        # cem_maxwell_init (src/cem_maxwell.F:65-191)
        self.cemface = np.zeros(nxzfl, dtype=np.int32)
        L.ora_set_fc_ptr(ldim, nx1, nelt, ip(self.cemface))
        self.pmltag = np.zeros(nelt, dtype=np.int32)
        self.pmlptr = np.zeros(max(nelt, 1), dtype=np.int32)
        self.maxpml = 0
        self.pmlinner = np.zeros(2 * ldim); self.pmlouter = np.zeros(2 * ldim)
        if self.ifpml:
            self.pmlthick = int(self.param.get(77, 1))
            self.pmlorder = float(self.param.get(78, 3.0))
            self.pmlreferr = float(self.param.get(79, 1e-6))
            faceary = self._pml_fill_faceary(self.pmlthick)
            self._pml_extent_and_tags(faceary)
        self.permittivity = z(npts); self.permeability = z(npts)
        if uservp is None:
            self.permittivity[:] = 1.0
            self.permeability[:] = 1.0
        else:
            uservp(self)
        assert self.permittivity.min() > 0 and self.permeability.min() > 0
        self._materials()
        self.pmlsigma = z(3 * npts)
        if self.ifpml:
            self._pml_calc_sigma()
        # cem_maxwell_pec_init (src/cem_maxwell.F:1338-1366)
        pec = []
        for e in range(nelt):
            for f in range(self.nfaces):
                if mesh.cbc[e][f] in ("PEC", "PML"):
                    base = e * self.nfaces * self.nxzf + f * self.nxzf
                    pec.extend(range(base, base + self.nxzf))
        self.cempec = np.array(pec if pec else [0], dtype=np.int32)
        self.ncempec = len(pec)
        # inverse mass incl. materials (src/cem_maxwell.F:183-186)
        self.ebm1 = 1.0 / (self.permittivity * self.bmn)
        self.hbm1 = 1.0 / (self.permeability * self.bmn)

        # fields
        self.hn, self.en = z(3 * npts), z(3 * npts)
        self.khn, self.ken = z(3 * npts), z(3 * npts)
        self.reshn, self.resen = z(3 * npts), z(3 * npts)
        self.fhn, self.fen = z(3 * nxzfl), z(3 * nxzfl)
        self.srflx = z(6 * nxzfl)
        self.pmlbn, self.pmldn = z(3 * npts), z(3 * npts)
        self.respmlbn, self.respmldn = z(3 * npts), z(3 * npts)
        self.respmlhn, self.respmlen = z(3 * npts), z(3 * npts)
        self.kpmlbn, self.kpmldn = z(3 * npts), z(3 * npts)

        # state struct
        s = OraState()
        s.ldim, s.nx1, s.nelt, s.imode = ldim, nx1, nelt, self.imode
        s.nxyz, s.nxzf, s.nfaces, s.npts, s.nxzfl = (self.nxyz, self.nxzf, self.nfaces, npts,
                                                      nxzfl)
        s.ifupwind, s.ifcentral = int(upwind), int(not upwind)
        s.ifpml, s.ifpec = int(self.ifpml), int(self.ifpec)
        for name in ("dxm1", "dxtm1", "w3mn", "rxmn", "rymn", "rzmn", "sxmn", "symn", "szmn",
                     "txmn", "tymn", "tzmn", "bmn", "unxm", "unym", "unzm", "aream", "hn",
                     "en", "khn", "ken", "reshn", "resen", "fhn", "fen", "srflx", "hbm1",
                     "ebm1", "Y_0", "Y_1", "Z_0", "Z_1", "permittivity", "permeability",
                     "pmlsigma", "pmlbn", "pmldn", "respmlbn", "respmldn", "respmlhn",
                     "respmlen", "kpmlbn", "kpmldn"):
            setattr(s, name, dp(getattr(self, name)))
        s.cemface, s.ncemface = ip(self.cemface), nxzfl
        s.cempec, s.ncempec = ip(self.cempec), self.ncempec
        s.gsh_face = self.gsh
        s.maxpml, s.pmlptr = self.maxpml, ip(self.pmlptr)
        s.time, s.dt, s.istep = 0.0, 0.0, 0
        self.s = s
        L.ora_rk_storage(C.byref(s))
        self._cbs = {}

    # -- helpers ---------------------------------------------------------------------
    def _arcsrf(self):
        """ARCSRF (src/nek5_genxyz.F:2-108, non-axisymmetric branch) for every edge flagged 'C',
        in GENXYZ's order (elements, then ISID = 1..8): the edge between preprocessor corners
        ISID and ISID+1 becomes a circular arc of radius CURVE(1,ISID,IE); the deviation from
        the straight edge is blended into the element with the linear hat functions (ADDTNSR,
        src/nek5_mat1.F:1003-1019).  libm sin/cos/atan2 point by point (math.*), so that the
        coordinates equal the reference's bit for bit."""
        import math
        mesh, n, ldim = self.mesh, self.nx1, self.ldim
        nz = n if ldim == 3 else 1
        z = self.zgm1
        h = [(1.0 - z) * 0.5, (1.0 + z) * 0.5]                 # H(.,d,1), H(.,d,2), d = 1,2
        h3 = h if ldim == 3 else [np.ones(nz), np.ones(nz)]     # H(.,3,.)
        eface1 = (3, 2, 4, 1)
        for e in range(self.nelt):
            sl = slice(e * self.nxyz, (e + 1) * self.nxyz)
            X = self.xm1[sl].reshape(nz, n, n)                  # [iz, iy, ix]
            Y = self.ym1[sl].reshape(nz, n, n)
            for isid in range(1, 9):
                if mesh.ccurve[e][isid - 1] != "C":
                    continue
                nxt = {4: 1, 8: 5}.get(isid, isid + 1)
                pt1x, pt1y = float(mesh.xc[e, isid - 1]), float(mesh.yc[e, isid - 1])
                pt2x, pt2y = float(mesh.xc[e, nxt - 1]), float(mesh.yc[e, nxt - 1])
                radius = float(mesh.curve[e, isid - 1, 0])
                gap = math.sqrt((pt1x - pt2x) ** 2 + (pt1y - pt2y) ** 2)
                if abs(2.0 * radius) <= gap * 1.00001:
                    raise ValueError("arcsrf: radius too small for side %d of element %d" % (isid, e + 1))
                xs = pt2y - pt1y
                ys = pt1x - pt2x
                xys = math.sqrt(xs ** 2 + ys ** 2)
                dtheta = abs(math.asin(0.5 * gap / radius))
                pt12x = (pt1x + pt2x) / 2.0
                pt12y = (pt1y + pt2y) / 2.0
                xcenn = pt12x - xs / xys * radius * math.cos(dtheta)
                ycenn = pt12y - ys / xys * radius * math.cos(dtheta)
                theta0 = math.atan2(pt12y - ycenn, pt12x - xcenn)
                isid1 = (isid - 1) % 4 + 1                       # MOD1(ISID,4)
                xcrv = np.zeros(n); ycrv = np.zeros(n)
                for ix in range(n):
                    ixt = n - 1 - ix if isid1 > 2 else ix
                    r = float(z[ix])
                    if radius < 0.0:
                        r = -r
                    xcrv[ixt] = (xcenn + abs(radius) * math.cos(theta0 + r * dtheta)
                                 - (float(h[0][ix]) * pt1x + float(h[1][ix]) * pt2x))
                    ycrv[ixt] = (ycenn + abs(radius) * math.sin(theta0 + r * dtheta)
                                 - (float(h[0][ix]) * pt1y + float(h[1][ix]) * pt2y))
                f = eface1[isid1 - 1]
                hz = h3[(isid - 1) // 4]
                if f <= 2:   # x-face: S(ix,iy,iz) += (crv(iy)*H3(iz)) * H(ix,1,f)
                    for S, crv in ((X, xcrv), (Y, ycrv)):
                        hh = crv[None, :] * hz[:, None]          # [iz, iy]
                        S += hh[:, :, None] * h[f - 1][None, None, :]
                else:        # y-face: S += (H(iy,2,f-2)*H3(iz)) * crv(ix)
                    for S, crv in ((X, xcrv), (Y, ycrv)):
                        hh = h[f - 3][None, :] * hz[:, None]
                        S += hh[:, :, None] * crv[None, None, :]

    def vtk_payload(self, which: str, as_double: bool = False) -> bytes:
        """payload of the VTK "VECTORS" block cem_out writes for EN ('en') / HN ('hn'):
        vtk_nonswap_field (src/io_dumpvtk.F:858-878) interleaves the components per node,
        writefield4 / writefield4_double (src/io_co.c:443-456, 511-524) cast to float / keep
        double and swap to big-endian"""
        a = (self.en if which == "en" else self.hn).reshape(3, self.npts)
        inter = np.ascontiguousarray(a.T)                       # (npts, 3): x,y,z per node
        return inter.astype(">f8" if as_double else ">f4").tobytes()

    def comp(self, arr, c):
        n = arr.size // 3
        return arr[c * n:(c + 1) * n]

    def set_callback(self, which: str, fn):
        """which in {'userinc','usersrc','userfsrc'}; fn(tt, a1..a6) with numpy views whose
        meaning/order is exactly the reference's argument list at that call site."""
        n = self.nxzfl if which in ("userinc", "userfsrc") else self.npts

        def tramp(tt, a1, a2, a3, a4, a5, a6, ctx):
            arrs = [np.ctypeslib.as_array(a, shape=(n,)) for a in (a1, a2, a3, a4, a5, a6)]
            fn(tt, *arrs)

        cb = USERCB(tramp)
        self._cbs[which] = cb
        setattr(self.s, which, cb)

    def set_dt(self, param12: float):
        """set_dt src/cem_drive.F:351-395 (Maxwell branch)."""
        if param12 < 0:
            self.s.dt = abs(param12)
        elif param12 > 0:
            dxmin = self.L.ora_get_dxmin(self.ldim, self.nx1, self.nelt, dp(self.xm1),
                                         dp(self.ym1), dp(self.zm1))
            self.s.dt = param12 * dxmin
        else:
            raise ValueError("set param(12) with nonzero")
        return self.s.dt

    @property
    def time(self):
        return self.s.time

    @property
    def dt(self):
        return self.s.dt

    def step(self, nsteps: int = 1):
        if getattr(self, "filter", None) is None:
            self.L.ora_advance(C.byref(self.s), nsteps)
            return
        for _ in range(nsteps):  # cem_maxwell_op_rk: `if (iffilter) call q_filter(0.01)` (:342)
            self.L.ora_advance(C.byref(self.s), 1)
            self.L.ora_q_filter(C.byref(self.s), dp(self.filter))

    def set_filter(self, wght: float = 0.01, ncut: int = 2):
        """param(18) = 1: iffilter (src/cem_param.F:71); q_filter builds its matrix once with
        ncut = 2 and the weight cem_maxwell_op_rk passes (0.01)"""
        self.filter = build_new_filter(self.zgm1, ncut, wght)

    def stage(self, rkstep: int):
        """One RK stage (rk_c; cem_maxwell_op; rk_maxwell_ab), 1-based rkstep."""
        s = self.s
        s.rkstep = rkstep
        s.rktime = s.time + s.dt * s.rk4c[rkstep - 1]
        self.L.ora_cem_maxwell_op(C.byref(s))
        self.L.ora_rk_maxwell_ab(C.byref(s), rkstep)

    def cem_error(self, u, exact):
        err = np.zeros(u.size)
        l2 = C.c_double(); linf = C.c_double()
        self.L.ora_cem_error(dp(np.ascontiguousarray(u)), dp(np.ascontiguousarray(exact)),
                             dp(err), u.size, dp(self.bmn), self.volvm1, C.byref(l2),
                             C.byref(linf))
        return l2.value, linf.value

    def errors(self, usersol):
        """userchk: six (l2, linf) pairs against usersol(time) -> (shn(3*npts), sen(3*npts))."""
        shn, sen = usersol(self, self.s.time)
        l2, linf = [], []
        for arr, sol in ((self.hn, shn), (self.en, sen)):
            for c in range(3):
                a, b = self.cem_error(self.comp(arr, c), self.comp(sol, c))
                l2.append(a); linf.append(b)
        return l2, linf

    # -- materials ---------------------------------------------------------------------
    def _materials(self):
        """cem_maxwell_materials src/cem_maxwell.F:262-325 (incl. the PEC doubling quirk)."""
        nxzfl = self.nxzfl
        impede = np.sqrt(self.permeability / self.permittivity)
        conduc = np.sqrt(self.permittivity / self.permeability)
        zimpede = impede[self.cemface].copy()
        yconduc = conduc[self.cemface].copy()
        Z_0 = zimpede.copy(); Y_0 = yconduc.copy()
        self.L.ora_gs_op_fields(self.gsh, dp(Z_0), nxzfl, 1, 1)
        self.L.ora_gs_op_fields(self.gsh, dp(Y_0), nxzfl, 1, 1)
        Y_1 = np.zeros(nxzfl); Z_1 = np.zeros(nxzfl)
        for e in range(self.nelt):
            for f in range(self.nfaces):
                if self.mesh.cbc[e][f] == "PEC":
                    b = e * self.nxzf * self.nfaces + self.nxzf * f
                    sl = slice(b, b + self.nxzf)
                    Y_0[sl] = 2.0 * Y_0[sl]; Y_1[sl] = 2.0 * Y_1[sl]
                    Z_0[sl] = 2.0 * Z_0[sl]; Z_1[sl] = 2.0 * Z_1[sl]
        Z_1 = Z_0 - zimpede
        Y_1 = Y_0 - yconduc
        Z_0 = 0.5 * Z_0
        Y_0 = 0.5 * Y_0
        self.Y_0, self.Y_1, self.Z_0, self.Z_1 = Y_0, Y_1, Z_0, Z_1
        self.yconduc = yconduc  # COMMON /EMWAVE/ yconduc: read by the graphene currents

    # -- PML setup -----------------------------------------------------------------------
    def _march_faces(self, faceary):
        """march_faces src/cem_maxwell_pml.F:85-137."""
        for axis in range(self.ldim):
            pos = EFACE[2 + 2 * axis - 1] - 1
            neg = EFACE[1 + 2 * axis - 1] - 1
            s = faceary[:, pos, :] + faceary[:, neg, :]
            m = s != 0
            faceary[:, pos, :][m] += 1
            faceary[:, neg, :][m] += 1
        flat = faceary.reshape(-1)
        self.L.ora_gs_op_fields(self.gsh, dp(flat), self.nxzfl, 1, 4)

    def _pml_fill_faceary(self, thick):
        """pml_fill_faceary src/cem_maxwell_pml.F:139-190."""
        n = self.nx1
        faceary = np.zeros((self.nelt, self.nfaces, self.nxzf))
        for e in range(self.nelt):
            for f in range(self.nfaces):
                if self.mesh.cbc[e][f] in ("PML", "pml"):
                    if self.ldim == 3:
                        for ix in range(1, n - 1):
                            for iz in range(1, n - 1):
                                faceary[e, f, ix + iz * n] = 1
                    else:
                        faceary[e, f, 1:n - 1] = 1
        for _ in range(thick):
            self._march_faces(faceary)
        return faceary

    def _dir_local_to_global(self, e, d):
        """dir_local_to_global src/cem_maxwell_pml.F:192-259 (d, result: sym faces 1..6)."""
        locvec = [0.0, 0.0, 0.0]
        axis = (d - 1) // 2
        locvec[axis] = -1.0 if (d - 1) % 2 == 0 else 1.0
        o = self.nxyz * e
        g = [self.rxmn[o] * locvec[0] + self.sxmn[o] * locvec[1] + self.txmn[o] * locvec[2],
             self.rymn[o] * locvec[0] + self.symn[o] * locvec[1] + self.tymn[o] * locvec[2],
             self.rzmn[o] * locvec[0] + self.szmn[o] * locvec[1] + self.tzmn[o] * locvec[2]]
        biggest, argmax = 0.0, 0
        for i in range(3):
            if abs(g[i]) >= biggest:
                biggest = abs(g[i]); argmax = i + 1
        globdir = (argmax - 1) * 2 + 1
        if g[argmax - 1] >= 0:
            globdir += 1
        return globdir

    def _pml_extent_and_tags(self, faceary):
        """pml_extent_and_tags src/cem_maxwell_pml.F:261-431."""
        ldim = self.ldim
        pmlinf = 1e20
        oppface = (2, 1, 4, 3, 6, 5)
        ind = (1 + self.nx1) if ldim == 3 else 1  # 0-based indicative point
        inner = np.zeros(2 * ldim); outer = np.zeros(2 * ldim)
        for axis in range(ldim):
            outer[2 * axis] = pmlinf; inner[2 * axis] = -pmlinf
            inner[2 * axis + 1] = pmlinf; outer[2 * axis + 1] = -pmlinf
        coords = (self.xm1, self.ym1, self.zm1)
        for e in range(self.nelt):
            tag = 0
            for axis in range(1, ldim + 1):
                face = (axis - 1) * 2 + 1
                far_here = faceary[e, EFACE[face - 1] - 1, ind]
                far_opp = faceary[e, EFACE[oppface[face - 1] - 1] - 1, ind]
                if far_here != 0 and far_opp != 0:
                    if far_here == far_opp:
                        raise RuntimeError('No "gradient" in PML indicators.')
                    globface = self._dir_local_to_global(e, face)
                    globaxis = (globface - 1) // 2 + 1
                    globsign = ((globface - 1) % 2) * 2 - 1
                    globface = (globaxis - 1) * 2 + 1
                    c = coords[globaxis - 1][self.nxyz * e:self.nxyz * (e + 1)]
                    mincoord, maxcoord = c.min(), c.max()
                    opp = oppface[globface - 1]
                    if globsign * (far_here - far_opp) > 0:
                        inner[opp - 1] = min(inner[opp - 1], mincoord)
                        outer[opp - 1] = max(outer[opp - 1], maxcoord)
                        tag |= 1 << (opp - 1)
                    if globsign * (far_here - far_opp) < 0:
                        inner[globface - 1] = max(inner[globface - 1], maxcoord)
                        outer[globface - 1] = min(outer[globface - 1], mincoord)
                        tag |= 1 << (globface - 1)
            self.pmltag[e] = tag
        self.pmlinner, self.pmlouter = inner, outer

    def _pml_calc_sigma(self):
        """pml_calc_sigma src/cem_maxwell_pml.F:433-506."""
        l = 0
        for e in range(self.nelt):
            if self.pmltag[e] != 0:
                self.pmlptr[l] = e
                l += 1
        self.maxpml = l
        order, referr = self.pmlorder, self.pmlreferr
        coords = (self.xm1, self.ym1, self.zm1)
        npts = self.npts
        for q in range(self.maxpml):
            e = int(self.pmlptr[q])
            sl = slice(self.nxyz * e, self.nxyz * (e + 1))
            for face in range(1, 2 * self.ldim + 1):
                axis = (face - 1) // 2 + 1
                width = abs(self.pmlouter[face - 1] - self.pmlinner[face - 1])
                if self.pmltag[e] & (1 << (face - 1)):
                    eta = np.sqrt(self.permeability[sl] / self.permittivity[sl])
                    sigmamax = -(order + 1) * math.log(referr) / (2 * eta * width)
                    zero2one = (coords[axis - 1][sl] - self.pmlinner[face - 1]) / (
                        self.pmlouter[face - 1] - self.pmlinner[face - 1])
                    # zero2one**order with a REAL exponent is libm pow() in the reference;
                    # numpy's vectorised power is not correctly rounded (1 ulp off in ~10 %
                    # of the points against oracle/_ref), so call libm element by element
                    z2o = np.array([math.pow(v, order) for v in zero2one])
                    self.pmlsigma[(axis - 1) * npts + sl.start:(axis - 1) * npts + sl.stop] = (
                        sigmamax * z2o)

    def sync_pml_state(self):
        """Refresh the struct after setup helpers changed counts."""
        self.s.maxpml = self.maxpml
