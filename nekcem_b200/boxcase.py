"""Synthetic periodic/PEC box cases: the arrays the Fortran host would hold in its COMMON
blocks after cem_init/cem_maxwell_init, generated directly for axis-aligned uniform boxes.

This is product-side host code (bench.py / smoke / the public API use it); it does NOT
import the oracle.  It plays the role of the reference's setup layer (L3) for
the one mesh family BASELINE.json's metric is quoted on: "synthetic 3D periodic box, E^3 hex
elements" with the tests/3dboxper initial condition.

Reference behaviour restated (no code copied):
  GLL nodes/weights, D matrix      src/nek5_speclib.F:107-122, 807-840
  element order, box vertices      makebox, src/nek5_genbox.F:402-560 (x fastest)
  pencil element->rank map         gfdm_elm_to_proc / gfdm_map_2d, src/nek5_map22.F:163-245
  metrics (unnormalised cofactors) GLMAPM1 src/nek5_coef.F:555-636, bm1 GEODAT1 :756-778
  face area / normals / slot order AREA3 src/nek5_coef.F:1150-1237 (-y,+x,+y,-x,-z,+z)
  materials (eps=mu=1) -> Y/Z      cem_maxwell_materials src/cem_maxwell.F:262-325
  initial condition / exact sol.   tests/3dboxper/3dboxper.usr:45-86
"""
from __future__ import annotations

import math

import numpy as np


# ----------------------------------------------------------------------------------------
def gll(n: int):
    """n GLL nodes and weights on [-1,1] (Newton on (1-x^2) P'_{n-1})."""
    N = n - 1
    if n == 2:
        return np.array([-1.0, 1.0]), np.array([1.0, 1.0])
    x = -np.cos(np.pi * np.arange(n) / N)
    for _ in range(100):
        P = np.zeros((n, n))
        P[:, 0] = 1.0
        P[:, 1] = x
        for k in range(2, n):
            P[:, k] = ((2 * k - 1) * x * P[:, k - 1] - (k - 1) * P[:, k - 2]) / k
        dx = (x * P[:, N] - P[:, N - 1]) / (n * P[:, N])
        x = x - dx
        if np.max(np.abs(dx)) < 1e-16:
            break
    x[0], x[-1] = -1.0, 1.0
    x = 0.5 * (x - x[::-1])  # enforce symmetry
    P = np.zeros((n, n))
    P[:, 0] = 1.0
    P[:, 1] = x
    for k in range(2, n):
        P[:, k] = ((2 * k - 1) * x * P[:, k - 1] - (k - 1) * P[:, k - 2]) / k
    w = 2.0 / (N * n * P[:, N] ** 2)
    return x, w


def legendre(x, N):
    p1, p2 = np.ones_like(x), x.copy()
    if N == 0:
        return p1
    for k in range(1, N):
        p1, p2 = p2, ((2 * k + 1) * x * p2 - k * p1) / (k + 1)
    return p2


def dgll(z: np.ndarray) -> np.ndarray:
    """D(i,j) = dl_j/dx(z_i) (DGLL), returned as an (n,n) array indexed [i,j]."""
    n = z.size
    N = n - 1
    L = legendre(z, N)
    D = np.zeros((n, n))
    for i in range(n):
        for j in range(n):
            if i != j:
                D[i, j] = L[i] / (L[j] * (z[i] - z[j]))
    D[0, 0] = -N * (N + 1) / 4.0
    D[N, N] = N * (N + 1) / 4.0
    return D


# ----------------------------------------------------------------------------------------
def gllnid_box(nelx: int, nely: int, nelz: int, nproc: int) -> np.ndarray:
    """Element -> rank map of the reference for .box (gtp) meshes: pencils along x dealt to
    ranks in a boustrophedon sweep of the (y,z) plane (gfdm_map_2d)."""
    nes, net = nely, nelz
    nep = nelx
    if nes * net < nproc:
        # the reference folds the x direction into the pencil plane here
        # (src/nek5_map22.F:187-191); not needed for one-box runs with <= 8 ranks
        raise NotImplementedError("more ranks than (y,z) pencils")
    num_el = [0] * (nproc + 1)
    k = nproc - 1
    for _ in range(nes * net):
        num_el[k] += 1
        k -= 1
        if k < 0:
            k = nproc - 1
    map_st = np.zeros((nes, net), dtype=np.int32)
    jnid, cnt, cur = 0, 0, num_el[0]
    for j in range(0, net, 2):
        for i in range(nes):
            cnt += 1
            if cnt > cur:
                jnid += 1
                cur = num_el[jnid]
                cnt = 1
            map_st[i, j] = jnid
        if j + 1 < net:
            for i in range(nes - 1, -1, -1):
                cnt += 1
                if cnt > cur:
                    jnid += 1
                    cur = num_el[jnid]
                    cnt = 1
                map_st[i, j + 1] = jnid
    g = np.zeros(nelx * nely * nelz, dtype=np.int32)
    ex = np.arange(nelx)
    for jt in range(nelz):
        for js in range(nely):
            g[ex + nelx * js + nelx * nely * jt] = map_st[js, jt]
    return g


def usersol_3dboxper(x, y, z, t):
    """tests/3dboxper/3dboxper.usr:64-80 (omega = sqrt(3) standing mode)."""
    om = math.sqrt(3.0)
    th, te = math.sin(om * t) / om, math.cos(om * t)
    return ((2 * np.cos(x) * np.sin(y) * np.cos(z) * th, -np.sin(x) * np.cos(y) * np.cos(z) * th,
             np.sin(x) * np.sin(y) * np.sin(z) * th),
            (0.0 * x * y * z, np.cos(x) * np.sin(y) * np.sin(z) * te,
             np.cos(x) * np.cos(y) * np.cos(z) * te))


class BoxCase:
    """Uniform periodic box [0,L]^3 of nel=(EX,EY,EZ) elements, order N = nx1-1, eps=mu=1.

    ``rank``/``nranks`` select the local elements by the reference's pencil map; local order is
    ascending global element id (lglel)."""

    def __init__(self, nel, nx1, rank=0, nranks=1, length=2 * math.pi, bc="P", eps_upper=1.0,
                 pml=False):
        """``eps_upper`` != 1: permittivity eps_upper in the upper half of the box (element rows
        ey >= EY/2), 1 below -- the two-material layout of tests/3ddielectric at benchmark size:
        masses, impedances Y_0..Z_1 as cem_maxwell_materials builds them (src/cem_maxwell.F:262-325).
        ``pml``: True / "all": every element is a PML element with a smooth synthetic sigma profile
        (the tests/3dboxpml layout at benchmark size; the numbers only have to keep the run
        bounded); "layers": the two bottom and two top element rows in y (the PML thickness of
        tests/3ddielectric)."""
        self.eps_upper = float(eps_upper)
        self.pml = "all" if pml is True else (pml or False)
        self.nel = tuple(int(v) for v in nel)
        self.nx1 = int(nx1)
        self.rank, self.nranks = rank, nranks
        self.length = float(length)
        self.bc = bc
        EX, EY, EZ = self.nel
        if bc == "P" and min(self.nel) < 3:
            raise ValueError("periodic directions need >= 3 elements (face ids from vertices)")
        self.gllnid = gllnid_box(EX, EY, EZ, nranks)
        self.lglel = np.nonzero(self.gllnid == rank)[0].astype(np.int64)
        self.nelt = int(self.lglel.size)
        n = self.nx1
        self.nxyz, self.nxzf, self.nfaces = n ** 3, n ** 2, 6
        self.npts = self.nxyz * self.nelt
        self.nxzfl = self.nxzf * 6 * self.nelt
        self.z, self.w = gll(n)
        self.D = dgll(self.z)
        self.h = (self.length / EX, self.length / EY, self.length / EZ)
        self.volume_global = self.length ** 3

    # element lattice coordinates of the local elements
    def _exyz(self):
        EX, EY, EZ = self.nel
        g = self.lglel
        return g % EX, (g // EX) % EY, g // (EX * EY)

    def coords(self):
        """xm1, ym1, zm1 (npts each) of the local elements."""
        n = self.nx1
        ex, ey, ez = self._exyz()
        hx, hy, hz = self.h
        xi = (self.z + 1.0) * 0.5
        X = (ex[:, None] + xi[None, :]) * hx  # (nelt, n)
        Y = (ey[:, None] + xi[None, :]) * hy
        Z = (ez[:, None] + xi[None, :]) * hz
        shape = (self.nelt, n, n, n)
        x = np.broadcast_to(X[:, None, None, :], shape).reshape(-1)
        y = np.broadcast_to(Y[:, None, :, None], shape).reshape(-1)
        zc = np.broadcast_to(Z[:, :, None, None], shape).reshape(-1)
        return x, y, zc

    def fields(self, t=0.0):
        """(hn, en), each (3*npts), from the 3dboxper exact solution at time t (separable
        evaluation: no npts-sized temporaries beyond the outputs)."""
        n = self.nx1
        ex, ey, ez = self._exyz()
        hx, hy, hz = self.h
        xi = (self.z + 1.0) * 0.5
        X = ((ex[:, None] + xi[None, :]) * hx)[:, None, None, :]
        Y = ((ey[:, None] + xi[None, :]) * hy)[:, None, :, None]
        Z = ((ez[:, None] + xi[None, :]) * hz)[:, :, None, None]
        cx, sx, cy, sy, cz, sz = np.cos(X), np.sin(X), np.cos(Y), np.sin(Y), np.cos(Z), np.sin(Z)
        om = math.sqrt(3.0)
        th, te = math.sin(om * t) / om, math.cos(om * t)
        hn = np.empty((3, self.nelt, n, n, n))
        en = np.empty((3, self.nelt, n, n, n))
        np.multiply(cx * (2 * th), sy * cz, out=hn[0])
        np.multiply(sx * (-th), cy * cz, out=hn[1])
        np.multiply(sx * th, sy * sz, out=hn[2])
        en[0] = 0.0
        np.multiply(cx * te, sy * sz, out=en[1])
        np.multiply(cx * te, cy * cz, out=en[2])
        return hn.reshape(-1), en.reshape(-1)

    def face_ids(self):
        """Face-point global ids (int64, nxzfl): the id of a face is that of the '+' side
        face of the element owning it in that direction; boundary faces (bc != 'P') get 0."""
        EX, EY, EZ = self.nel
        NE = EX * EY * EZ
        n2 = self.nxzf
        ex, ey, ez = self._exyz()
        g = self.lglel
        per = self.bc == "P"

        def nbr(dx, dy, dz):
            return ((ex + dx) % EX) + EX * (((ey + dy) % EY) + EY * ((ez + dz) % EZ))

        p = np.arange(n2, dtype=np.int64)[None, :]
        out = np.zeros((self.nelt, 6, n2), dtype=np.int64)
        # slot order -y,+x,+y,-x,-z,+z ; direction index 0=x,1=y,2=z
        spec = [(0, 1, nbr(0, -1, 0), ey == 0), (1, 0, g, ex == EX - 1),
                (2, 1, g, ey == EY - 1), (3, 0, nbr(-1, 0, 0), ex == 0),
                (4, 2, nbr(0, 0, -1), ez == 0), (5, 2, g, ez == EZ - 1)]
        for slot, d, owner, onb in spec:
            fid = (d * NE + owner.astype(np.int64))[:, None]
            ids = fid * n2 + p + 1
            if not per:
                ids = np.where(onb[:, None], 0, ids)
            out[:, slot, :] = ids
        return out.reshape(-1)

    ARRAY_NAMES = ("dxm1", "w3mn", "rxmn", "rymn", "rzmn", "sxmn", "symn", "szmn", "txmn",
                   "tymn", "tzmn", "bmn", "hbm1", "ebm1", "unxm", "unym", "unzm", "aream",
                   "Y_0", "Y_1", "Z_0", "Z_1", "glo_num", "cempec", "pmlptr", "volvm1", "hn",
                   "en", "permittivity", "permeability", "pmlsigma", "pmlbn", "pmldn")

    def _eps_el(self):
        """permittivity of the local elements"""
        EX, EY, EZ = self.nel
        _, ey, _ = self._exyz()
        return np.where(ey >= EY // 2, self.eps_upper, 1.0)

    def array(self, name: str, t: float = 0.0):
        """One COMMON array by its reference name (generated on demand so that a 64^3 case
        never holds more than a couple of npts-sized host arrays at a time)."""
        n, nelt, npts, nxzfl = self.nx1, self.nelt, self.npts, self.nxzfl
        hx, hy, hz = self.h
        if name == "dxm1":
            return np.ascontiguousarray(self.D.T).reshape(-1)  # column-major: D(i,m) at i+n*m
        w3 = (self.w[None, None, :] * self.w[None, :, None] * self.w[:, None, None]).reshape(-1)
        if name == "w3mn":
            return w3
        if name == "rxmn":
            return np.full(npts, hy * hz / 4.0)
        if name == "symn":
            return np.full(npts, hx * hz / 4.0)
        if name == "tzmn":
            return np.full(npts, hx * hy / 4.0)
        if name in ("rymn", "rzmn", "sxmn", "szmn", "txmn", "tymn"):
            return np.zeros(npts)
        if name == "bmn":
            return np.tile((hx * hy * hz / 8.0) * w3, nelt)
        if name == "permeability":
            return np.ones(npts)
        if name == "permittivity":
            return np.repeat(self._eps_el(), self.nxyz)
        if name == "hbm1":  # 1/(mu*bm), mu = 1
            return np.tile(1.0 / ((hx * hy * hz / 8.0) * w3), nelt)
        if name == "ebm1":  # 1/(eps*bm)  (src/cem_maxwell.F:183-186)
            if self.eps_upper == 1.0:
                return np.tile(1.0 / ((hx * hy * hz / 8.0) * w3), nelt)
            bm = (hx * hy * hz / 8.0) * w3
            return (1.0 / (self._eps_el()[:, None] * bm[None, :])).reshape(-1)
        if name == "pmlsigma":
            # (npts,3): smooth, positive, below the stability bound of the synthetic dt
            x, y, zc = self.coords()
            L = self.length
            return np.concatenate([0.3 * np.sin(np.pi * x / L) ** 2, 0.3 * np.sin(np.pi * y / L) ** 2,
                                   0.3 * np.sin(np.pi * zc / L) ** 2])
        if name in ("pmlbn", "pmldn"):
            # B = mu H, D = eps E at t = 0 (userini of tests/3ddielectric, :67-78)
            hn, en = self.fields(t)
            if name == "pmlbn":
                return hn
            return (en.reshape(3, -1) * np.repeat(self._eps_el(), self.nxyz)[None, :]).reshape(-1)
        if name in ("unxm", "unym", "unzm", "aream"):
            f = np.zeros((6, self.nxzf))
            if name == "unxm":
                f[1], f[3] = 1.0, -1.0
            elif name == "unym":
                f[0], f[2] = -1.0, 1.0
            elif name == "unzm":
                f[4], f[5] = -1.0, 1.0
            else:
                ww = (self.w[None, :] * self.w[:, None]).reshape(-1)  # w(a)*w(b) at a + n*b
                f[0] = f[2] = (hx * hz / 4.0) * ww
                f[1] = f[3] = (hy * hz / 4.0) * ww
                f[4] = f[5] = (hx * hy / 4.0) * ww
            return np.tile(f.reshape(-1), nelt)
        if name in ("Y_0", "Y_1", "Z_0", "Z_1"):
            # eps = mu = 1: Z = Y = 1 on both sides; on PEC faces the materials quirk also
            # ends with Z_0 = Z_1 = Z^- (src/cem_maxwell.F:297-320)
            if self.eps_upper == 1.0:
                return np.ones(nxzfl)
            # two materials: own-side value + the neighbour's across each face;
            # X_0 = (X^- + X^+)/2, X_1 = X^+  (src/cem_maxwell.F:283-320)
            own = np.sqrt(self._eps_el()) if name[0] == "Y" else 1.0 / np.sqrt(self._eps_el())
            EX, EY, EZ = self.nel
            ex, ey, ez = self._exyz()
            up = lambda eyy: np.where(eyy % EY >= EY // 2, self.eps_upper, 1.0)
            nb_eps = np.stack([up(ey - 1), up(ey), up(ey + 1), up(ey), up(ey), up(ey)], axis=1)
            nbr = np.sqrt(nb_eps) if name[0] == "Y" else 1.0 / np.sqrt(nb_eps)
            val = 0.5 * (own[:, None] + nbr) if name[2] == "0" else nbr
            return np.repeat(val.reshape(-1), self.nxzf)
        if name == "glo_num":
            return self.face_ids()
        if name == "cempec":
            if self.bc == "P":
                return np.zeros(0, dtype=np.int64)
            return np.nonzero(self.face_ids() == 0)[0]
        if name == "pmlptr":
            if self.pml == "layers":
                _, ey, _ = self._exyz()
                return np.nonzero((ey < 2) | (ey >= self.nel[1] - 2))[0].astype(np.int64)
            return np.arange(nelt, dtype=np.int64) if self.pml else np.zeros(0, dtype=np.int64)
        if name == "volvm1":
            return self.volume_global
        if name in ("hn", "en"):
            hn, en = self.fields(t)
            return hn if name == "hn" else en
        raise KeyError(name)

    def lazy(self, with_fields=True, t=0.0):
        """dict-like view for MaxwellB200.cem_maxwell_init that generates arrays on access."""
        case = self

        class _Lazy(dict):
            def __contains__(self, k):
                if k in ("permittivity", "permeability", "pmlsigma", "pmlbn", "pmldn") and not case.pml:
                    return False
                return k in case.ARRAY_NAMES and (with_fields or k not in ("hn", "en"))

            def __getitem__(self, k):
                if k not in self:
                    raise KeyError(k)
                if k in ("hn", "en"):
                    if "_f" not in self.__dict__:
                        self.__dict__["_f"] = case.fields(t)
                    f = self.__dict__["_f"]
                    return f[0] if k == "hn" else f[1]
                return case.array(k, t)

            def __setitem__(self, k, v):  # cem_maxwell_init(free_after_upload=True) hook
                if k in ("hn", "en") and v is None:
                    self.__dict__.pop("_f", None)

            def get(self, k, default=None):
                return self[k] if k in self else default

        return _Lazy()

    def arrays(self, with_fields=True, t=0.0) -> dict:
        """All COMMON arrays for MaxwellB200.cem_maxwell_init as a plain dict."""
        aux = ("permittivity", "permeability", "pmlsigma", "pmlbn", "pmldn")
        return {k: self.array(k, t) for k in self.ARRAY_NAMES
                if (with_fields or k not in ("hn", "en")) and (self.pml or k not in aux)}


class BoxCase2D:
    """Uniform periodic box [0,L]^2 of nel=(EX,EY) elements, order N = nx1-1, eps=mu=1, for the 2D
    TE (imode 1: Ex,Ey,Hz) / TM (imode 2: Hx,Hy,Ez) path: the arrays of tests/2dboxper at any size,
    generated analytically (conventions as the reference's 2D setup: rxm1 = hy/2, sym1 = hx/2,
    tzm1 = 1, faces -y,+x,+y,-x).  Single rank."""

    ARRAY_NAMES = BoxCase.ARRAY_NAMES[:28]

    def __init__(self, nel, nx1, imode=1, length=2 * math.pi):
        self.nel = tuple(int(v) for v in nel)
        self.nx1, self.imode, self.length = int(nx1), int(imode), float(length)
        EX, EY = self.nel
        if min(self.nel) < 3:
            raise ValueError("periodic directions need >= 3 elements")
        n = self.nx1
        self.nelt = EX * EY
        self.nxyz, self.nxzf, self.nfaces = n * n, n, 4
        self.npts = self.nxyz * self.nelt
        self.nxzfl = self.nxzf * 4 * self.nelt
        self.z, self.w = gll(n)
        self.D = dgll(self.z)
        self.h = (self.length / EX, self.length / EY)
        self.volume_global = self.length ** 2
        self.lglel = np.arange(self.nelt, dtype=np.int64)

    def coords(self):
        n = self.nx1
        EX, EY = self.nel
        g = np.arange(self.nelt)
        ex, ey = g % EX, g // EX
        xi = (self.z + 1.0) * 0.5
        X = (ex[:, None] + xi[None, :]) * self.h[0]
        Y = (ey[:, None] + xi[None, :]) * self.h[1]
        shape = (self.nelt, n, n)
        return (np.broadcast_to(X[:, None, :], shape).reshape(-1),
                np.broadcast_to(Y[:, :, None], shape).reshape(-1))

    def fields(self, t=0.0):
        """(hn, en) of tests/2dboxper/2dboxper.usr:33-105 (omega = sqrt(2))"""
        x, y = self.coords()
        n = self.npts
        om = math.sqrt(2.0)
        hn = np.zeros(3 * n); en = np.zeros(3 * n)
        if self.imode == 2:
            th, te = math.sin(om * t) / om, math.cos(om * t)
            hn[0:n] = np.cos(x) * np.sin(y) * th
            hn[n:2 * n] = -np.sin(x) * np.cos(y) * th
            en[2 * n:] = np.cos(x) * np.cos(y) * te
        else:
            th, te = math.cos(om * t), math.sin(om * t) / om
            hn[2 * n:] = np.sin(x) * np.sin(y) * th
            en[0:n] = np.sin(x) * np.cos(y) * te
            en[n:2 * n] = -np.cos(x) * np.sin(y) * te
        return hn, en

    def face_ids(self):
        EX, EY = self.nel
        NE = EX * EY
        n = self.nx1
        g = np.arange(NE)
        ex, ey = g % EX, g // EX
        nbr = lambda dx, dy: ((ex + dx) % EX) + EX * ((ey + dy) % EY)
        p = np.arange(n, dtype=np.int64)[None, :]
        out = np.zeros((NE, 4, n), dtype=np.int64)
        for slot, d, owner in ((0, 1, nbr(0, -1)), (1, 0, g), (2, 1, g), (3, 0, nbr(-1, 0))):
            out[:, slot, :] = (d * NE + owner.astype(np.int64))[:, None] * n + p + 1
        return out.reshape(-1)

    def array(self, name, t=0.0):
        n, nelt, npts, nxzfl = self.nx1, self.nelt, self.npts, self.nxzfl
        hx, hy = self.h
        if name == "dxm1":
            return np.ascontiguousarray(self.D.T).reshape(-1)
        w3 = (self.w[None, :] * self.w[:, None]).reshape(-1)
        if name == "w3mn":
            return w3
        if name == "rxmn":
            return np.full(npts, hy / 2.0)
        if name == "symn":
            return np.full(npts, hx / 2.0)
        if name == "tzmn":
            return np.ones(npts)
        if name in ("rymn", "rzmn", "sxmn", "szmn", "txmn", "tymn"):
            return np.zeros(npts)
        if name == "bmn":
            return np.tile((hx * hy / 4.0) * w3, nelt)
        if name in ("hbm1", "ebm1"):
            return np.tile(1.0 / ((hx * hy / 4.0) * w3), nelt)
        if name in ("unxm", "unym", "unzm", "aream"):
            f = np.zeros((4, n))
            if name == "unxm":
                f[1], f[3] = 1.0, -1.0
            elif name == "unym":
                f[0], f[2] = -1.0, 1.0
            elif name == "aream":
                f[0] = f[2] = (hx / 2.0) * self.w
                f[1] = f[3] = (hy / 2.0) * self.w
            return np.tile(f.reshape(-1), nelt)
        if name in ("Y_0", "Y_1", "Z_0", "Z_1"):
            return np.ones(nxzfl)
        if name == "glo_num":
            return self.face_ids()
        if name in ("cempec", "pmlptr"):
            return np.zeros(0, dtype=np.int64)
        if name == "volvm1":
            return self.volume_global
        if name in ("hn", "en"):
            hn, en = self.fields(t)
            return hn if name == "hn" else en
        raise KeyError(name)

    def arrays(self, t=0.0):
        return {k: self.array(k, t) for k in self.ARRAY_NAMES}
