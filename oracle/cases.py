"""Reference test cases restated for the oracle (test infrastructure, NOT product code).

Each builder returns a ``RefCase`` with the reference's .usr callbacks (usrdat2, uservp,
userini, userinc, usersrc, usersol) restated in numpy, plus the tolerances the reference's
``userchk`` enforces.  File:line citations point at /root/reference/tests/<case>/.
"""
from __future__ import annotations

import math
import os

import numpy as np

from . import oracle as O

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "tests", "golden")


def _load_mesh(name):
    d = np.load(os.path.join(GOLDEN, name))
    extra = dict(ccurve=d["ccurve"], curve=d["curve"]) if "ccurve" in d.files else {}
    return O.mesh_from_arrays(int(d["ndim"]), d["xc"], d["yc"], d["zc"],
                              [[str(c) for c in row] for row in d["cbc"]], d["vertex"],
                              **extra), d["part"]


def _rescale(case, lo, hi):
    """usrdat2 pattern: affine map of the mesh bounding box onto [lo,hi]^d."""
    for arr, l, h in zip((case.xm1, case.ym1, case.zm1)[:case.ldim], lo, hi):
        mn, mx = arr.min(), arr.max()
        s = (h - l) / (mx - mn)
        arr[:] = s * (arr - mn) + l


# ------------------------------------------------------------------------------------
# tests/3dboxper : periodic box, standing mode omega = sqrt(3)
# ------------------------------------------------------------------------------------
def usersol_3dboxper(case, tt):
    """tests/3dboxper/3dboxper.usr:45-86"""
    omega = math.sqrt(3.0)
    tmph = math.sin(omega * tt) / omega
    tmpe = math.cos(omega * tt)
    xx, yy, zz = case.xm1, case.ym1, case.zm1
    n = case.npts
    shn = np.zeros(3 * n); sen = np.zeros(3 * n)
    shn[0:n] = 2 * np.cos(xx) * np.sin(yy) * np.cos(zz) * tmph
    shn[n:2 * n] = -np.sin(xx) * np.cos(yy) * np.cos(zz) * tmph
    shn[2 * n:] = np.sin(xx) * np.sin(yy) * np.sin(zz) * tmph
    sen[0:n] = 0.0
    sen[n:2 * n] = np.cos(xx) * np.sin(yy) * np.sin(zz) * tmpe
    sen[2 * n:] = np.cos(xx) * np.cos(yy) * np.cos(zz) * tmpe
    return shn, sen


def usrdat2_3dboxper(case):
    """tests/3dboxper/3dboxper.usr:138-167: sx*(x-xmin) with sx = 2*pi/(xmax-xmin)."""
    pi = 4.0 * math.atan(1.0)
    for arr in (case.xm1, case.ym1, case.zm1):
        mn, mx = arr.min(), arr.max()
        s = 2 * pi / (mx - mn)
        arr[:] = s * (arr - mn)


def case_3dboxper(nx1=9, mesh=None):
    """tests/3dboxper (128 hex from .re2/.map, all periodic; N=8; dt=2e-4, 50 steps, upwind).
    Tolerances: L2 5e-10, Linf 5e-9 (3dboxper.usr:181-193)."""
    if mesh is None:
        mesh, _ = _load_mesh("3dboxper_mesh.npz")
    c = O.RefCase(mesh, nx1, upwind=True, usrdat2=usrdat2_3dboxper)
    c.set_dt(-0.0002)
    c.usersol = usersol_3dboxper
    shn, sen = usersol_3dboxper(c, 0.0)
    c.hn[:] = shn; c.en[:] = sen
    c.tol = dict(l2=[5e-10] * 6, linf=[5e-9] * 6)
    c.nsteps = 50
    return c


def case_boxper(nel=(2, 2, 2), nx1=6, dt=-0.0002):
    """Synthetic periodic .box variant of 3dboxper on [0,2pi]^3."""
    pi2 = 2 * 4.0 * math.atan(1.0)
    mesh = O.box_mesh(nel, ((0.0, pi2),) * 3, ("P  ",) * 6)
    c = O.RefCase(mesh, nx1, upwind=True)
    c.set_dt(dt)
    c.usersol = usersol_3dboxper
    shn, sen = usersol_3dboxper(c, 0.0)
    c.hn[:] = shn; c.en[:] = sen
    return c


# ------------------------------------------------------------------------------------
# tests/3dboxpec : PEC cavity
# ------------------------------------------------------------------------------------
def usersol_3dboxpec(case, tt):
    """tests/3dboxpec/3dboxpec.usr:64-115"""
    pi = 4.0 * math.atan(1.0)
    ww = pi
    sqrt2, sqrt3, sqrt6 = math.sqrt(2.0), math.sqrt(3.0), math.sqrt(6.0)
    tmph = math.cos(ww * sqrt3 * tt) / sqrt6
    tmpe = math.sin(ww * sqrt3 * tt) / sqrt2
    xx, yy, zz = case.xm1, case.ym1, case.zm1
    n = case.npts
    shn = np.zeros(3 * n); sen = np.zeros(3 * n)
    shn[0:n] = -np.sin(ww * xx) * np.cos(ww * yy) * np.cos(ww * zz) * tmph
    shn[n:2 * n] = -np.cos(ww * xx) * np.sin(ww * yy) * np.cos(ww * zz) * tmph
    shn[2 * n:] = 2 * np.cos(ww * xx) * np.cos(ww * yy) * np.sin(ww * zz) * tmph
    sen[0:n] = -np.cos(ww * xx) * np.sin(ww * yy) * np.sin(ww * zz) * tmpe
    sen[n:2 * n] = np.sin(ww * xx) * np.cos(ww * yy) * np.sin(ww * zz) * tmpe
    sen[2 * n:] = 0
    return shn, sen


def usrdat2_3dboxpec(case):
    """tests/3dboxpec/3dboxpec.usr:151-183: sx*(x-xmin)-1 with sx = 2/(xmax-xmin)."""
    for arr in (case.xm1, case.ym1, case.zm1):
        mn, mx = arr.min(), arr.max()
        s = 2.0 / (mx - mn)
        arr[:] = s * (arr - mn) - 1


def case_3dboxpec(nx1=9, mesh=None):
    """tests/3dboxpec (27 hex from .rea/.map, PEC walls; N=8; dt=5e-3, 1000 steps).
    Tolerances: L2 5e-8, Linf 5e-7 (3dboxpec.usr:198-210)."""
    if mesh is None:
        mesh, _ = _load_mesh("3dboxpec_mesh.npz")
    c = O.RefCase(mesh, nx1, upwind=True, usrdat2=usrdat2_3dboxpec)
    c.set_dt(-0.005)
    c.usersol = usersol_3dboxpec
    shn, sen = usersol_3dboxpec(c, 0.0)
    c.hn[:] = shn; c.en[:] = sen
    c.tol = dict(l2=[5e-8] * 6, linf=[5e-7] * 6)
    c.nsteps = 1000
    return c


def case_boxpec(nel=(2, 2, 2), nx1=6, dt=-0.005):
    """Synthetic PEC .box cavity on [-1,1]^3 with the 3dboxpec solution."""
    mesh = O.box_mesh(nel, ((-1.0, 1.0),) * 3, ("PEC",) * 6)
    c = O.RefCase(mesh, nx1, upwind=True)
    c.set_dt(dt)
    c.usersol = usersol_3dboxpec
    shn, sen = usersol_3dboxpec(c, 0.0)
    c.hn[:] = shn; c.en[:] = sen
    return c


# ------------------------------------------------------------------------------------
# tests/3ddielectric : plane wave through a dielectric interface, PML in +-y
# ------------------------------------------------------------------------------------
class _Dielectric:
    """tests/3ddielectric/3ddielectric.usr (uservp :170-262, usrdat2 :271-303,
    userinc :6-50, userini :52-81, usersol :83-151)."""

    def __init__(self, twomat: bool):
        self.twomat = twomat
        self.omega = 2.0
        self.eps1 = 1.0
        self.eps2 = 2.0 if twomat else 1.0
        self.mu1 = self.mu2 = 1.0
        z1 = math.sqrt(self.mu1 / self.eps1)
        z2 = math.sqrt(self.mu2 / self.eps2)
        self.reflte = (z1 - z2) / (z1 + z2)
        self.trante = 2 * z1 / (z1 + z2)
        self.refltm = (z2 - z1) / (z1 + z2)
        self.trantm = 2 * z2 / (z1 + z2)

    def usrdat2(self, case):
        for arr, s in zip((case.xm1, case.ym1, case.zm1), (5.0, 10.0, 5.0)):
            mn, mx = arr.min(), arr.max()
            arr[:] = s * (arr - mn) / (mx - mn) - (s / 2.0)

    def _mid(self, case):
        # ym1(nx1/2,nx1/2,nx1/2,e), 1-based integer division
        h = case.nx1 // 2 - 1
        n = case.nx1
        return h + n * h + n * n * h

    def uservp(self, case):
        ym = case.ym1.reshape(case.nelt, case.nxyz)
        upper = ym[:, self._mid(case)] > 0
        eps = np.where(upper, self.eps1, self.eps2)
        mu = np.where(upper, self.mu1, self.mu2)
        case.permittivity[:] = np.repeat(eps, case.nxyz)
        case.permeability[:] = np.repeat(mu, case.nxyz)
        self.upper = upper
        inc = []
        for e in range(case.nelt):
            if upper[e]:
                # NB: the reference never resets `markinc` to .true. per face
                # (3ddielectric.usr:233-252): once a face fails, later faces of that
                # element are skipped.  Only the -y face (slot 1, checked first) can
                # qualify on this mesh, so the quirk is harmless but restated.
                markinc = True
                for f in range(case.nfaces):
                    base = e * case.nxzf * case.nfaces + case.nxzf * f
                    js = np.arange(base, base + case.nxzf)
                    ks = case.cemface[js]
                    if np.any(np.abs(case.ym1[ks]) > 1e-8):
                        markinc = False
                    if markinc:
                        inc.extend(js.tolist())
        self.incindex = np.array(inc, dtype=np.int64)

    def userinc(self, case):
        j = self.incindex
        k = case.cemface[j]
        eps = case.permittivity[k]; mu = case.permeability[k]
        eta = np.sqrt(mu / eps)
        ky = self.omega * np.sqrt(mu * eps)
        yy = case.ym1[k]

        def cb(tt, fhx, fhy, fhz, fex, fey, fez):
            uinc = np.cos(-ky * yy - self.omega * tt)
            fhz[j] = fhz[j] + uinc
            fex[j] = fex[j] + eta * uinc
            fez[j] = fez[j] + uinc
            fhx[j] = fhx[j] - uinc / eta

        return cb

    def usersol(self, case, tt):
        n = case.npts
        eps = case.permittivity; mu = case.permeability
        eta = np.sqrt(mu / eps)
        ky = self.omega * np.sqrt(eps * mu)
        yy = case.ym1
        upper = np.repeat(self.upper, case.nxyz)
        inpml = np.repeat(case.pmltag != 0, case.nxyz)
        order, referr = case.pmlorder, case.pmlreferr
        pmlfac = np.zeros(n)
        # upper region (+y PML: sym face 4)
        d = case.pmlouter[3] - case.pmlinner[3]
        smax = -(order + 1) * math.log(referr) / (2 * eta * d)
        with np.errstate(invalid="ignore"):
            fu = (smax * d / (order + 1)) * ((yy - case.pmlinner[3]) / d) ** (order + 1)
        d2 = case.pmlinner[2] - case.pmlouter[2]
        smax2 = -(order + 1) * math.log(referr) / (2 * eta * d2)
        with np.errstate(invalid="ignore"):
            fl = (smax2 * d2 / (order + 1)) * ((case.pmlinner[2] - yy) / d2) ** (order + 1)
        pmlfac = np.where(inpml, np.where(upper, fu, fl), 0.0)
        uu_u = np.exp(-eta * pmlfac) * np.cos(ky * yy - self.omega * tt)
        uu_l = np.exp(-eta * pmlfac) * np.cos(-ky * yy - self.omega * tt)
        shn = np.zeros(3 * n); sen = np.zeros(3 * n)
        shn[2 * n:] = np.where(upper, self.reflte * uu_u, self.trante * uu_l)          # hz
        sen[0:n] = np.where(upper, -eta * self.reflte * uu_u, eta * self.trante * uu_l)  # ex
        sen[2 * n:] = np.where(upper, self.refltm * uu_u, self.trantm * uu_l)          # ez
        shn[0:n] = np.where(upper, self.refltm * uu_u / eta, -self.trantm * uu_l / eta)  # hx
        return shn, sen


def case_3ddielectric(twomat=False, nx1=9, nel=(4, 8, 4)):
    """tests/3ddielectric (.box 4x8x4, BC P,P,PML,PML,P,P; N=8; dt=5e-3; 1000 steps;
    param(70)=1 -> two materials; PML thick 2, order 3, referr 1e-10).
    Tolerances 5e-4 / 5e-3 on hx,hz,ex,ez (3ddielectric.usr userchk)."""
    mesh = O.box_mesh(nel, ((-1.0, 1.0),) * 3, ("P  ", "P  ", "PML", "PML", "P  ", "P  "))
    u = _Dielectric(twomat)
    c = O.RefCase(mesh, nx1, upwind=True, usrdat2=u.usrdat2, uservp=u.uservp,
                  param={77: 2, 78: 3.0, 79: 1e-10, 70: 1 if twomat else 0})
    c.user = u
    c.set_dt(-0.005)
    c.usersol = lambda case, tt: u.usersol(case, tt)
    shn, sen = u.usersol(c, 0.0)
    c.hn[:] = shn; c.en[:] = sen
    n = c.npts
    for k in range(3):  # userini :67-78
        c.pmlbn[k * n:(k + 1) * n] = c.permeability * shn[k * n:(k + 1) * n]
        c.pmldn[k * n:(k + 1) * n] = c.permittivity * sen[k * n:(k + 1) * n]
    c.set_callback("userinc", u.userinc(c))
    c.tol = dict(l2=[5e-4] * 6, linf=[5e-3] * 6)
    c.nsteps = 1000
    return c


class _Dielectric2D(_Dielectric):
    """tests/2ddielectric/2ddielectric.usr: the same interface problem in 2D, one polarisation
    per run (TE: hz, ex; TM: ez, hx), omega = 5."""

    def __init__(self, imode: int, twomat: bool):
        super().__init__(twomat)
        self.imode = imode
        self.omega = 5.0
        z1 = math.sqrt(self.mu1 / self.eps1)
        z2 = math.sqrt(self.mu2 / self.eps2)
        if imode == 1:   # 2ddielectric.usr:241-243
            self.refl, self.tran = (z1 - z2) / (z1 + z2), 2 * z1 / (z1 + z2)
        else:            # :244-246
            self.refl, self.tran = (z2 - z1) / (z1 + z2), 2 * z2 / (z1 + z2)
        self.reflte = self.refltm = self.trante = self.trantm = None  # 3D names: unused

    def usrdat2(self, case):
        for arr, s in zip((case.xm1, case.ym1), (5.0, 10.0)):
            mn, mx = arr.min(), arr.max()
            arr[:] = s * (arr - mn) / (mx - mn) - (s / 2.0)

    def _mid(self, case):
        h = case.nx1 // 2 - 1
        return h + case.nx1 * h

    def _amp(self, eta):
        amp = np.zeros((6,) + np.shape(eta))
        if self.imode == 1:
            amp[2] = 1.0; amp[3] = eta
        else:
            amp[5] = 1.0; amp[0] = -1.0 / eta
        return amp

    def userinc(self, case):
        j = self.incindex
        k = case.cemface[j]
        eps = case.permittivity[k]; mu = case.permeability[k]
        eta = np.sqrt(mu / eps)
        ky = self.omega * np.sqrt(mu * eps)
        yy = case.ym1[k]

        def cb(tt, fhx, fhy, fhz, fex, fey, fez):
            uinc = np.cos(-ky * yy - self.omega * tt)
            if self.imode == 1:
                fhz[j] = fhz[j] + uinc
                fex[j] = fex[j] + eta * uinc
            else:
                fez[j] = fez[j] + uinc
                fhx[j] = fhx[j] - uinc / eta

        return cb

    def incident(self, case):
        """the same userinc as arguments of MaxwellB200.set_incident"""
        j = self.incindex
        k = case.cemface[j]
        eps = case.permittivity[k]; mu = case.permeability[k]
        eta = np.sqrt(mu / eps)
        ky = self.omega * np.sqrt(mu * eps)
        return j, self._amp(eta), -ky * case.ym1[k], self.omega

    def usersol(self, case, tt):
        n = case.npts
        eps = case.permittivity; mu = case.permeability
        eta = np.sqrt(mu / eps)
        ky = self.omega * np.sqrt(eps * mu)
        yy = case.ym1
        upper = np.repeat(self.upper, case.nxyz)
        inpml = np.repeat(case.pmltag != 0, case.nxyz)
        order, referr = case.pmlorder, case.pmlreferr
        d = case.pmlouter[3] - case.pmlinner[3]
        smax = -(order + 1) * math.log(referr) / (2 * eta * d)
        with np.errstate(invalid="ignore"):
            fu = (smax * d / (order + 1)) * ((yy - case.pmlinner[3]) / d) ** (order + 1)
        d2 = case.pmlinner[2] - case.pmlouter[2]
        smax2 = -(order + 1) * math.log(referr) / (2 * eta * d2)
        with np.errstate(invalid="ignore"):
            fl = (smax2 * d2 / (order + 1)) * ((case.pmlinner[2] - yy) / d2) ** (order + 1)
        pmlfac = np.where(inpml, np.where(upper, fu, fl), 0.0)
        uu_u = self.refl * np.exp(-eta * pmlfac) * np.cos(ky * yy - self.omega * tt)
        uu_l = self.tran * np.exp(-eta * pmlfac) * np.cos(-ky * yy - self.omega * tt)
        shn = np.zeros(3 * n); sen = np.zeros(3 * n)
        if self.imode == 1:
            shn[2 * n:] = np.where(upper, uu_u, uu_l)                # hz
            sen[0:n] = np.where(upper, -eta * uu_u, eta * uu_l)      # ex
        else:
            sen[2 * n:] = np.where(upper, uu_u, uu_l)                # ez
            shn[0:n] = np.where(upper, uu_u / eta, -uu_l / eta)      # hx
        return shn, sen


def case_2ddielectric(imode=1, twomat=False, nx1=9, nel=(4, 32)):
    """tests/2ddielectric (.box 4x32, BC P,P,PML,PML; N=8; dt=5e-3; 1000 steps; param(70)=1 ->
    two materials; PML thick 10, order 3, referr 1e-30).  Tolerances (2ddielectric.usr userchk):
    one material 5e-8 / 1e-6 (TM Linf 5e-6), two materials 5e-7 / 5e-6; zero component
    ~1e-14 / 1e-12."""
    mesh = O.box_mesh(nel, ((-1500.0, 1500.0),) * 2, ("P  ", "P  ", "PML", "PML"))
    u = _Dielectric2D(imode, twomat)
    c = O.RefCase(mesh, nx1, imode=imode, upwind=True, usrdat2=u.usrdat2, uservp=u.uservp,
                  param={77: 10, 78: 3.0, 79: 1e-30, 70: 1 if twomat else 0})
    c.user = u
    c.set_dt(-0.005)
    c.usersol = lambda case, tt: u.usersol(case, tt)
    shn, sen = u.usersol(c, 0.0)
    c.hn[:] = shn; c.en[:] = sen
    n = c.npts
    for k in range(3):  # userini
        c.pmlbn[k * n:(k + 1) * n] = c.permeability * shn[k * n:(k + 1) * n]
        c.pmldn[k * n:(k + 1) * n] = c.permittivity * sen[k * n:(k + 1) * n]
    c.set_callback("userinc", u.userinc(c))
    if twomat:
        w, z, wi, zi = 5e-7, 5e-15 if imode == 1 else 1e-14, 5e-6, 5e-13
    else:
        w, z, wi, zi = 5e-8, 1e-14, 1e-6 if imode == 1 else 5e-6, 1e-12
    if imode == 1:   # hz, ex wave; ey zero
        c.tol = dict(l2=[0, 0, w, w, z, 0], linf=[0, 0, wi, wi, zi, 0])
    else:            # hx, ez wave; hy zero
        c.tol = dict(l2=[w, z, 0, 0, 0, w], linf=[wi, zi, 0, 0, 0, wi])
    c.nsteps = 1000
    return c


# ------------------------------------------------------------------------------------
# tests/3dboxpml : Gaussian-pulsed dipole in an all-PML box (stability check only)
# ------------------------------------------------------------------------------------
def usersrc_3dboxpml(case):
    """tests/3dboxpml/3dboxpml.usr:30-88: Gaussian approximation of a Hertzian dipole,
    srcez -= i0*norm*exp(-r^2/(2 w^2)) * sin(-omega t) * bm1."""
    omega, width = 2.0, 0.1
    norm = 1.0 / (math.sqrt(8 * math.atan(1.0)) * width) ** 3
    i0 = 1.0 / norm
    xfac = -0.5 * ((case.xm1 - 0.0) / width) ** 2
    yfac = -0.5 * ((case.ym1 - 0.0) / width) ** 2
    zfac = -0.5 * ((case.zm1 - 0.0) / width) ** 2
    g = i0 * norm * np.exp(xfac + yfac + zfac)

    def cb(tt, shx, shy, shz, sex, sey, sez):
        tfac = math.sin(-omega * tt)
        sez[:] = sez - g * (tfac * case.bmn)

    cb.profile = g  # spatial profile (per node), time factor sin(-omega t) * bm1
    cb.omega = omega
    return cb


def case_3dboxpml(nx1=9, nel=(6, 6, 6)):
    """tests/3dboxpml (.box 6^3, all PML, thick 1, order 3, referr 1e-8, CFL 0.1; zero
    initial fields; usrdat2 maps onto [-1,1]^3).  userchk only bounds |fields| <= 1."""
    mesh = O.box_mesh(nel, ((-1.0, 1.0),) * 3, ("PML",) * 6)

    def usrdat2(case):
        for arr in (case.xm1, case.ym1, case.zm1):
            mn, mx = arr.min(), arr.max()
            arr[:] = 2.0 * (arr - mn) / (mx - mn) - 2.0 / 2.0

    c = O.RefCase(mesh, nx1, upwind=True, usrdat2=usrdat2, param={77: 1, 78: 3.0, 79: 1e-8})
    c.set_dt(0.1)
    c.usersrc_fn = usersrc_3dboxpml(c)
    c.set_callback("usersrc", c.usersrc_fn)
    c.usersol = lambda case, tt: (np.zeros(3 * case.npts), np.zeros(3 * case.npts))
    c.tol = dict(l2=[1.0] * 6, linf=[1.0] * 6)
    c.nsteps = 2000
    return c


def usersrc_2dboxpml(case):
    """tests/2dboxpml/2dboxpml.usr usersrc: 2D Gaussian source, srchz (TE) or srcez (TM)
    -= i0*norm*exp(-r^2/(2 w^2)) * sin(-omega t) * bm1 with norm = 1/(2 pi w^2)."""
    omega, width, i0 = 2.0, 0.1, 1.0
    norm = 1.0 / (8 * math.atan(1.0) * width ** 2)
    xfac = -0.5 * ((case.xm1 - 0.0) / width) ** 2
    yfac = -0.5 * ((case.ym1 - 0.0) / width) ** 2
    g = i0 * norm * np.exp(xfac + yfac)
    te = case.imode == 1

    def cb(tt, shx, shy, shz, sex, sey, sez):
        tfac = math.sin(-omega * tt)
        tgt = shz if te else sez
        tgt[:] = tgt - g * (tfac * case.bmn)

    cb.profile = g
    cb.omega = omega
    cb.comp = 2 if te else 5
    return cb


def case_2dboxpml(imode=1, nx1=9, nel=(8, 8)):
    """tests/2dboxpml (.box 8x8, all PML, thick 2, order 3, referr 1e-8, CFL 0.1; zero initial
    fields; usrdat2 maps onto [-1,1]^2; 4000 steps).  userchk only bounds |fields| <= 1."""
    mesh = O.box_mesh(nel, ((-1500.0, 1500.0),) * 2, ("PML",) * 4)

    def usrdat2(case):
        for arr in (case.xm1, case.ym1):
            mn, mx = arr.min(), arr.max()
            arr[:] = 2.0 * (arr - mn) / (mx - mn) - 2.0 / 2.0

    c = O.RefCase(mesh, nx1, imode=imode, upwind=True, usrdat2=usrdat2,
                  param={77: 2, 78: 3.0, 79: 1e-8})
    c.set_dt(0.1)
    c.usersrc_fn = usersrc_2dboxpml(c)
    c.set_callback("usersrc", c.usersrc_fn)
    c.usersol = lambda case, tt: (np.zeros(3 * case.npts), np.zeros(3 * case.npts))
    c.tol = dict(l2=[1.0] * 6, linf=[1.0] * 6)
    c.nsteps = 4000
    return c


# ------------------------------------------------------------------------------------
# tests/2dboxper, tests/2dboxpec : 2D TE / TM standing modes
# ------------------------------------------------------------------------------------
def usersol_2dboxper(case, tt):
    """tests/2dboxper/2dboxper.usr:33-105 (omega = sqrt(2); TM: hx,hy,ez; TE: hz,ex,ey)."""
    n = case.npts
    xx, yy = case.xm1, case.ym1
    omega = math.sqrt(2.0)
    shn = np.zeros(3 * n); sen = np.zeros(3 * n)
    if case.imode == 2:
        tmph = math.sin(omega * tt) / omega; tmpe = math.cos(omega * tt)
        shn[0:n] = np.cos(xx) * np.sin(yy) * tmph
        shn[n:2 * n] = -np.sin(xx) * np.cos(yy) * tmph
        sen[2 * n:] = np.cos(xx) * np.cos(yy) * tmpe
    else:
        tmph = math.cos(omega * tt); tmpe = math.sin(omega * tt) / omega
        shn[2 * n:] = np.sin(xx) * np.sin(yy) * tmph
        sen[0:n] = np.sin(xx) * np.cos(yy) * tmpe
        sen[n:2 * n] = -np.cos(xx) * np.sin(yy) * tmpe
    return shn, sen


def case_2dboxper(imode=1, nx1=9, nel=(3, 3), dt=-0.005):
    """tests/2dboxper (9 elements, N=8, dt=5e-3, 1000 steps; usrdat2 :155-187 maps onto
    [0,2pi]^2; userchk tolerances 5e-8 / 5e-7 on the three active components, :199-231)."""
    mesh = O.box_mesh(nel, ((0.0, 1.0),) * 2, ("P  ",) * 4)
    c = O.RefCase(mesh, nx1, imode=imode, upwind=True,
                  usrdat2=lambda case: _rescale(case, (0.0, 0.0), (2 * math.pi, 2 * math.pi)))
    c.set_dt(dt)
    c.usersol = usersol_2dboxper
    shn, sen = usersol_2dboxper(c, 0.0)
    c.hn[:] = shn; c.en[:] = sen
    act = [0, 0, 1, 1, 1, 0] if imode == 1 else [1, 1, 0, 0, 0, 1]
    c.tol = dict(l2=[5e-8 * a for a in act], linf=[5e-7 * a for a in act])
    c.nsteps = 1000
    return c


def usersol_2dboxpec(case, tt):
    """tests/2dboxpec/2dboxpec.usr:40-108 (ww = 1.5 pi, omega = sqrt(2))."""
    n = case.npts
    xx, yy = case.xm1, case.ym1
    ww = 1.5 * math.pi
    omega = math.sqrt(2.0)
    tmph = math.sin(ww * omega * tt) / omega
    tmpe = math.cos(ww * omega * tt)
    shn = np.zeros(3 * n); sen = np.zeros(3 * n)
    if case.imode == 2:
        shn[0:n] = np.cos(ww * xx) * np.sin(ww * yy) * tmph
        shn[n:2 * n] = -np.sin(ww * xx) * np.cos(ww * yy) * tmph
        sen[2 * n:] = np.cos(ww * xx) * np.cos(ww * yy) * tmpe
    else:
        shn[2 * n:] = np.sin(ww * xx) * np.sin(ww * yy) * tmpe
        sen[0:n] = np.sin(ww * xx) * np.cos(ww * yy) * tmph
        sen[n:2 * n] = -np.cos(ww * xx) * np.sin(ww * yy) * tmph
    return shn, sen


def case_2dboxpec(imode=1, nx1=9, nel=(3, 3), dt=-0.005):
    """tests/2dboxpec (9 elements, PEC walls, usrdat2 :176-187 maps onto [-1,1]^2; tolerances
    1e-6 / 1e-5, :204-231)."""
    mesh = O.box_mesh(nel, ((0.0, 1.0),) * 2, ("PEC",) * 4)
    c = O.RefCase(mesh, nx1, imode=imode, upwind=True,
                  usrdat2=lambda case: _rescale(case, (-1.0, -1.0), (1.0, 1.0)))
    c.set_dt(dt)
    c.usersol = usersol_2dboxpec
    shn, sen = usersol_2dboxpec(c, 0.0)
    c.hn[:] = shn; c.en[:] = sen
    act = [0, 0, 1, 1, 1, 0] if imode == 1 else [1, 1, 0, 0, 0, 1]
    c.tol = dict(l2=[1e-6 * a for a in act], linf=[1e-5 * a for a in act])
    c.nsteps = 1000
    return c


# ------------------------------------------------------------------------------------
# tests/drude, tests/lorentz : 2D TE plane wave onto a dispersive half space, ADE for J
# ------------------------------------------------------------------------------------
class _Dispersive:
    """tests/drude/drude.usr and tests/lorentz/lorentz.usr (uservp, usrdat2, userinc, userini,
    usersol, usersrc).  kind = 'drude' | 'lorentz'."""

    def __init__(self, kind: str):
        self.kind = kind
        self.omega = 5.0
        self.mu1 = self.mu2 = 1.0
        self.eps1 = 1.0
        om = self.omega
        if kind == "drude":
            self.pa, self.pb = 0.0, 100.0  # mydrudea, mydrudeb (drude.usr:236-237)
            self.eps2 = 1.0 - self.pb / (om * (om + 1j * self.pa))
        else:
            self.pa = 0.0
            self.pb = 0.9 * om ** 2
            self.pc = 1.0 * self.pb  # lorentz.usr:242-244
            self.eps2 = 1.0 + self.pc / (self.pb - 1j * self.pa * om - om ** 2)
        import cmath
        self.eta1 = math.sqrt(self.mu1 / self.eps1)
        eta2 = cmath.sqrt(self.mu2 / self.eps2)
        if eta2.real == 0.0 and eta2.imag > 0.0:  # branch fix (drude.usr:243-247)
            eta2 = -eta2
        self.eta2 = eta2
        self.k1 = om * math.sqrt(self.mu1 * self.eps1)
        k2 = om * cmath.sqrt(self.mu2 * self.eps2)
        if k2.imag < 0:  # Fortran's csqrt(-x + 0i) = +i sqrt(x): the decaying branch
            k2 = -k2
        self.k2 = k2
        self.refl = (self.eta1 - eta2) / (self.eta1 + eta2)
        self.tran = 2 * self.eta1 / (self.eta1 + eta2)

    def usrdat2(self, case):
        for arr, s in zip((case.xm1, case.ym1), (5.0, 10.0)):
            mn, mx = arr.min(), arr.max()
            arr[:] = s * (arr - mn) / (mx - mn) - (s / 2.0)

    def _mid(self, case):
        h = case.nx1 // 2 - 1
        return h + case.nx1 * h

    def uservp(self, case):
        ym = case.ym1.reshape(case.nelt, case.nxyz)
        upper = ym[:, self._mid(case)] > 0
        self.upper = upper
        case.permittivity[:] = np.repeat(np.where(upper, self.eps1, 1.0), case.nxyz)
        case.permeability[:] = np.repeat(np.where(upper, self.mu1, self.mu2), case.nxyz)
        n = case.npts
        low = np.repeat(~upper, case.nxyz)
        self.index = np.nonzero(low)[0].astype(np.int32)  # 0-based node list
        npar = 2 if self.kind == "drude" else 3
        self.params = np.zeros(npar * n)
        self.params[0:n][low] = self.pa
        self.params[n:2 * n][low] = self.pb
        if npar == 3:
            self.params[2 * n:][low] = self.pc
        nj = 3 if self.kind == "drude" else 6
        self.jn = np.zeros(nj * n); self.kjn = np.zeros(nj * n); self.resjn = np.zeros(nj * n)
        inc = []
        for e in range(case.nelt):
            if upper[e]:
                markinc = True  # never reset per face, like the reference (drude.usr:289-305)
                for f in range(case.nfaces):
                    base = e * case.nxzf * case.nfaces + case.nxzf * f
                    js = np.arange(base, base + case.nxzf)
                    if np.any(np.abs(case.ym1[case.cemface[js]]) > 1e-8):
                        markinc = False
                    if markinc:
                        inc.extend(js.tolist())
        self.incindex = np.array(inc, dtype=np.int64)

    def userinc(self, case):
        j = self.incindex
        yy = case.ym1[case.cemface[j]]

        def cb(tt, fhx, fhy, fhz, fex, fey, fez):
            uinc = np.cos(-self.k1 * yy - self.omega * tt)
            fhz[j] = fhz[j] + uinc
            fex[j] = fex[j] + self.eta1 * uinc

        return cb

    def incident(self, case):
        """the same userinc as arguments of MaxwellB200.set_incident"""
        j = self.incindex
        yy = case.ym1[case.cemface[j]]
        amp = np.zeros((6, j.size))
        amp[2] = 1.0
        amp[3] = self.eta1
        return j, amp, -self.k1 * yy, self.omega

    def usersol(self, case, tt):
        n = case.npts
        yy = case.ym1
        upper = np.repeat(self.upper, case.nxyz)
        inpml = np.repeat(case.pmltag != 0, case.nxyz)
        d = case.pmlouter[3] - case.pmlinner[3]
        smax = -(case.pmlorder + 1) * math.log(case.pmlreferr) / (2.0 * self.eta1 * d)
        with np.errstate(invalid="ignore"):
            pf = (smax * d / (case.pmlorder + 1)) * ((yy - case.pmlinner[3]) / d) ** (case.pmlorder + 1)
        pmlfac = np.where(inpml & upper, pf, 0.0)
        usc_u = self.refl * np.exp(1j * (self.k1 * yy - self.omega * tt) - self.eta1 * pmlfac)
        usc_l = self.tran * np.exp(1j * (-self.k2 * yy - self.omega * tt))
        shn = np.zeros(3 * n); sen = np.zeros(3 * n)
        shn[2 * n:] = np.where(upper, usc_u.real, usc_l.real)
        sen[0:n] = np.where(upper, (-self.eta1 * usc_u).real, (self.eta2 * usc_l).real)
        return shn, sen

    def userini(self, case):
        n = case.npts
        shn, sen = self.usersol(case, 0.0)
        case.hn[:] = shn; case.en[:] = sen
        for k in range(3):
            case.pmlbn[k * n:(k + 1) * n] = case.permeability * shn[k * n:(k + 1) * n]
            case.pmldn[k * n:(k + 1) * n] = case.permittivity * sen[k * n:(k + 1) * n]
        j = self.index
        yy = case.ym1[j]
        efac = self.tran * self.eta2 * np.exp(1j * (-self.k2 * yy - self.omega * 0.0))
        if self.kind == "drude":
            sigma = self.params[n:2 * n][j] / (self.params[0:n][j] - 1j * self.omega)
            self.jn[0:n][j] = (sigma * efac).real
        else:
            a, b, c = self.params[0:n][j], self.params[n:2 * n][j], self.params[2 * n:][j]
            sigma = 1j * c * self.omega / (self.omega ** 2 + 1j * a * self.omega - b)
            jf = sigma * efac
            self.jn[0:n][j] = jf.real
            self.jn[3 * n:4 * n][j] = ((1j / self.omega) * jf).real

    def usersrc(self, case):
        fn = (case.L.ora_cem_maxwell_drude if self.kind == "drude"
              else case.L.ora_cem_maxwell_lorentz)
        import ctypes as C

        def cb(tt, *res):
            fn(C.byref(case.s), O.dp(self.jn), O.dp(self.kjn), O.dp(self.resjn),
               O.dp(self.params), O.ip(self.index), int(self.index.size))

        return cb


def _case_dispersive(kind, nx1, nel, thick):
    mesh = O.box_mesh(nel, ((-1500.0, 1500.0),) * 2, ("P  ", "P  ", "PEC", "PML"))
    u = _Dispersive(kind)
    c = O.RefCase(mesh, nx1, imode=1, upwind=True, usrdat2=u.usrdat2, uservp=u.uservp,
                  param={77: thick, 78: 3.0, 79: 1e-15})
    c.user = u
    c.set_dt(-0.005)
    c.usersol = lambda case, tt: u.usersol(case, tt)
    u.userini(c)
    c.set_callback("userinc", u.userinc(c))
    c.set_callback("usersrc", u.usersrc(c))
    c.nsteps = 1000
    return c


def case_drude(nx1=9, nel=(4, 32), thick=6):
    """tests/drude (.box 4x32, BC P,P,PEC,PML; N=8; dt=5e-3; 1000 steps; PML thick 6, order 3,
    referr 1e-15).  Tolerances 1e-7 / 5e-6 on hz, ex, ey (drude.usr userchk)."""
    c = _case_dispersive("drude", nx1, nel, thick)
    c.tol = dict(l2=[0, 0, 1e-7, 1e-7, 1e-7, 0], linf=[0, 0, 5e-6, 5e-6, 5e-6, 0])
    return c


def case_lorentz(nx1=9, nel=(4, 32), thick=6):
    """tests/lorentz (same mesh as drude).  Tolerances 5e-6 / 5e-5 on hz, ex; 5e-10 on ey
    (lorentz.usr:373-387)."""
    c = _case_dispersive("lorentz", nx1, nel, thick)
    c.tol = dict(l2=[0, 0, 5e-6, 5e-6, 5e-10, 0], linf=[0, 0, 5e-5, 5e-5, 5e-10, 0])
    return c


# ------------------------------------------------------------------------------------
# tests/cylwave : TM_01 mode of a circular PEC waveguide, periodic along z -- an unstructured
# mesh of 5 x 10 hexahedra whose outer sides are circular arcs (curved, non-affine elements)
# ------------------------------------------------------------------------------------
BSSLJ_RT_0_1 = 2.4048255576957729   # bsrt(0,1): first zero of J_0 (src/cem_bessel.F:870)


def geom_xyradius(mesh):
    """src/cem_common.F:817-843: largest radius of the element corners in the x-y plane"""
    return math.sqrt(float(np.max(mesh.xc * mesh.xc + mesh.yc * mesh.yc)))


def usrdat_cylwave(mesh):
    """cylwave.usr usrdat: corners within 1e-4 (squared) of the outer radius, rounded to one
    decimal, are projected onto it"""
    radius = geom_xyradius(mesh)
    radius = int(10 * radius + 0.1) / 10.0
    e1 = radius * radius - 1e-4
    rr = mesh.xc * mesh.xc + mesh.yc * mesh.yc
    outer = rr > e1
    rn = np.where(outer, radius / np.sqrt(np.where(outer, rr, 1.0)), 1.0)
    mesh.xc[:] = np.where(outer, rn * mesh.xc, mesh.xc)
    mesh.yc[:] = np.where(outer, rn * mesh.yc, mesh.yc)


def usrdat2_cylwave(case):
    """cylwave.usr usrdat2: the z extent becomes 2 pi xmax"""
    pi = 4.0 * math.atan(1.0)
    zmin, zmax = case.zm1.min(), case.zm1.max()
    sz = 2 * pi * case.xm1.max() / (zmax - zmin)
    case.zm1[:] = sz * (case.zm1 - zmin)


def usersol_cylwave(case, tt):
    """cylwave.usr usersol (de Wolf, Essentials of Electromagnetics for Engineering, 19.7): TM
    mode m = 0, first root, one wavelength along z.  J_0 and J_0' = -J_1 from scipy (the
    reference evaluates them with Cody's RJBESL, src/cem_bessel.F; not translated)."""
    from scipy import special
    n = case.npts
    pi = 4.0 * math.atan(1.0)
    radius = case.cyl_radius
    xx, yy, zz = case.xm1, case.ym1, case.zm1
    zsize = zz.max() - zz.min()
    kz = 2 * pi * 1.0 / zsize
    rho = np.sqrt(xx ** 2 + yy ** 2)
    phi = np.arctan2(yy, xx)
    bigk = BSSLJ_RT_0_1 / radius
    omega = math.sqrt(kz ** 2 + bigk ** 2)
    allfac = np.exp(-1j * kz * zz) * np.exp(1j * omega * tt)
    j0 = special.j0(bigk * rho)
    j0p = -special.j1(bigk * rho)
    ezz = (allfac * j0).real
    erho = (allfac * (-1j) * kz / bigk * j0p).real
    hphi = (allfac * (-1j) * omega * 1.0 / bigk * j0p).real
    shn = np.zeros(3 * n); sen = np.zeros(3 * n)
    shn[0:n] = -np.sin(phi) * hphi
    shn[n:2 * n] = np.cos(phi) * hphi
    sen[0:n] = np.cos(phi) * erho
    sen[n:2 * n] = np.sin(phi) * erho
    sen[2 * n:] = ezz
    return shn, sen


def case_cylwave(nx1=12):
    """tests/cylwave (50 elements from the reference's .rea/.map, 80 circular-arc sides, PEC wall,
    periodic in z; N=11; param(12)=+0.25 -> CFL dt; 1000 steps).  Tolerances 5e-9 / 5e-8 on all
    six components (cylwave.usr userchk)."""
    mesh, _ = _load_mesh("cylwave_mesh.npz")
    usrdat_cylwave(mesh)
    c = O.RefCase(mesh, nx1, upwind=True, usrdat2=usrdat2_cylwave)
    c.cyl_radius = geom_xyradius(mesh)
    c.set_dt(0.25)
    c.usersol = usersol_cylwave
    shn, sen = usersol_cylwave(c, 0.0)
    c.hn[:] = shn; c.en[:] = sen
    c.tol = dict(l2=[5e-9] * 6, linf=[5e-8] * 6)
    c.nsteps = 1000
    return c


# ------------------------------------------------------------------------------------
# tests/3dgraphene, tests/2dgraphene : plane wave(s) onto a flat graphene sheet at y = 0; the
# sheet is a surface current advanced by face-point ADEs inside userfsrc
# ------------------------------------------------------------------------------------
class _Graphene:
    """tests/3dgraphene/3dgraphene.usr and tests/2dgraphene/2dgraphene.usr (uservp, usrdat2,
    userinc, userini, usersol, userfsrc).  imode: 3 = 3D (TE and TM waves superimposed),
    1 = 2D TE, 2 = 2D TM."""

    # graphene parameters (3dgraphene.usr:326-337): Drude term + two critical-point terms
    PARAMS = (0.000e+00, 1.499e+00, -2.599e-03, 4.632e+05, 1.090e+03, -1.391e+00, -3.125e+02,
              -1.049e-03, 4.271e+05, 2.742e+02, -7.769e-02, 4.268e+02)

    def __init__(self, imode: int):
        self.imode = imode
        self.omega = om = 5.0
        self.eps1 = self.eps2 = self.mu1 = self.mu2 = 1.0
        (a_d, b_d, b_cp1, a_211, a_221, b_11, b_21, b_cp2, a_212, a_222, b_12,
         b_22) = self.PARAMS
        CI = 1j
        csigma_d = b_d / (a_d - CI * om)
        csigma_cp1 = (CI / om) * ((a_211 * b_11 + CI * om * b_21)
                                  / (om ** 2 - a_211 + CI * om * a_221) + b_11) - b_cp1
        csigma_cp2 = (CI / om) * ((a_212 * b_12 + CI * om * b_22)
                                  / (om ** 2 - a_212 + CI * om * a_222) + b_12) - b_cp2
        self.sigmagraph = sg = csigma_d + csigma_cp1 + csigma_cp2
        z1 = math.sqrt(self.mu1 / self.eps1)
        z2 = math.sqrt(self.mu2 / self.eps2)
        self.z1 = z1
        self.reflte = (z1 - z2 + sg * z1 * z2) / (z1 + z2 + sg * z1 * z2)
        self.trante = 2 * z1 / (z1 + z2 + sg * z1 * z2)
        self.refltm = (z2 - z1 - z1 * z2 * sg) / (z1 + z2 + z1 * z2 * sg)
        self.trantm = 2 * z2 / (z1 + z2 + z1 * z2 * sg)
        self.te = imode in (3, 1)  # which of the two polarisations are present
        self.tm = imode in (3, 2)

    def usrdat2(self, case):
        arrs = (case.xm1, case.ym1, case.zm1) if case.ldim == 3 else (case.xm1, case.ym1)
        for arr, s in zip(arrs, (5.0, 10.0, 5.0)):
            mn, mx = arr.min(), arr.max()
            arr[:] = s * (arr - mn) / (mx - mn) - (s / 2.0)

    def _mid(self, case):
        h = case.nx1 // 2 - 1
        n = case.nx1
        return h + n * h + (n * n * h if case.ldim == 3 else 0)

    def uservp(self, case):
        ym = case.ym1.reshape(case.nelt, case.nxyz)
        upper = ym[:, self._mid(case)] > 0
        self.upper = upper
        case.permittivity[:] = np.repeat(np.where(upper, self.eps1, self.eps2), case.nxyz)
        case.permeability[:] = np.repeat(np.where(upper, self.mu1, self.mu2), case.nxyz)
        inc, gr = [], []
        for e in range(case.nelt):
            markinc = True  # 2D: set once per element (2dgraphene.usr:397-398)
            for f in range(case.nfaces):
                base = e * case.nxzf * case.nfaces + case.nxzf * f
                js = np.arange(base, base + case.nxzf)
                onsheet = not np.any(np.abs(case.ym1[case.cemface[js]]) > 1e-8)
                if upper[e]:
                    if case.ldim == 3:
                        markinc = True  # 3D: reset per face (3dgraphene.usr:370)
                    if not onsheet:
                        markinc = False
                    if markinc:
                        inc.extend(js.tolist())
                if onsheet:
                    gr.extend(js.tolist())
        self.incindex = np.array(inc, dtype=np.int64)
        self.graphindex = np.array(gr, dtype=np.int32)  # 0-based face points
        nf = case.nxzfl
        self.graphparams = np.zeros(12 * nf)
        for q, v in enumerate(self.PARAMS):
            self.graphparams[q * nf + self.graphindex] = v
        self.fjn = np.zeros(18 * nf); self.kfjn = np.zeros(18 * nf); self.resfjn = np.zeros(18 * nf)

    def _inc_amp(self, eta):
        """amplitudes of userinc per component (hx,hy,hz,ex,ey,ez) for a unit uinc"""
        amp = np.zeros((6,) + np.shape(eta))
        if self.te:
            amp[2] = 1.0; amp[3] = eta
        if self.tm:
            amp[5] = 1.0; amp[0] = -1.0 / eta
        return amp

    def userinc(self, case):
        j = self.incindex
        k = case.cemface[j]
        eps = case.permittivity[k]; mu = case.permeability[k]
        eta = np.sqrt(mu / eps)
        ky = self.omega * np.sqrt(mu * eps)
        yy = case.ym1[k]

        def cb(tt, fhx, fhy, fhz, fex, fey, fez):
            uinc = np.cos(-ky * yy - self.omega * tt)
            if self.te:
                fhz[j] = fhz[j] + uinc
                fex[j] = fex[j] + eta * uinc
            if self.tm:
                fez[j] = fez[j] + uinc
                fhx[j] = fhx[j] - uinc / eta

        return cb

    def incident(self, case):
        """the same userinc as arguments of MaxwellB200.set_incident"""
        j = self.incindex
        k = case.cemface[j]
        eps = case.permittivity[k]; mu = case.permeability[k]
        eta = np.sqrt(mu / eps)
        ky = self.omega * np.sqrt(mu * eps)
        return j, self._inc_amp(eta), -ky * case.ym1[k], self.omega

    def usersol(self, case, tt):
        n = case.npts
        eps = case.permittivity; mu = case.permeability
        eta = np.sqrt(mu / eps)
        ky = self.omega * np.sqrt(eps * mu)
        yy = case.ym1
        upper = np.repeat(self.upper, case.nxyz)
        inpml = np.repeat(case.pmltag != 0, case.nxyz)
        order, referr = case.pmlorder, case.pmlreferr
        d = case.pmlouter[3] - case.pmlinner[3]
        smax = -(order + 1) * math.log(referr) / (2 * eta * d)
        with np.errstate(invalid="ignore"):
            fu = (smax * d / (order + 1)) * ((yy - case.pmlinner[3]) / d) ** (order + 1)
        d2 = case.pmlinner[2] - case.pmlouter[2]
        smax2 = -(order + 1) * math.log(referr) / (2 * eta * d2)
        with np.errstate(invalid="ignore"):
            fl = (smax2 * d2 / (order + 1)) * ((case.pmlinner[2] - yy) / d2) ** (order + 1)
        pmlfac = np.where(inpml, np.where(upper, fu, fl), 0.0)
        uu_u = np.exp(1j * (ky * yy - self.omega * tt) - eta * pmlfac)
        uu_l = np.exp(1j * (-ky * yy - self.omega * tt) - eta * pmlfac)
        shn = np.zeros(3 * n); sen = np.zeros(3 * n)
        if self.te:
            r, t = self.reflte, self.trante
            shn[2 * n:] = np.where(upper, (r * uu_u).real, (t * uu_l).real)                # hz
            sen[0:n] = np.where(upper, -(r * eta * uu_u).real, (t * eta * uu_l).real)      # ex
        if self.tm:
            r, t = self.refltm, self.trantm
            sen[2 * n:] = np.where(upper, (r * uu_u).real, (t * uu_l).real)                # ez
            shn[0:n] = np.where(upper, (r * uu_u / eta).real, -(t * uu_l / eta).real)      # hx
        return shn, sen

    def userini(self, case):
        n = case.npts
        shn, sen = self.usersol(case, 0.0)
        case.hn[:] = shn; case.en[:] = sen
        for k in range(3):
            case.pmlbn[k * n:(k + 1) * n] = case.permeability * shn[k * n:(k + 1) * n]
            case.pmldn[k * n:(k + 1) * n] = case.permittivity * sen[k * n:(k + 1) * n]
        # currents: 1/2 of the parallel part of the complex E field at the interface
        enpar = [0.5 * self.z1 * (1.0 - self.reflte) if self.te else 0.0, 0.0,
                 0.5 * (1.0 + self.refltm) if self.tm else 0.0]
        nf = case.nxzfl
        j = self.graphindex
        P = lambda q: self.graphparams[q * nf + j]
        om = self.omega
        CI = 1j

        def put(c, q, val):  # fjn(j, c+1, q+1)
            self.fjn[(c + 3 * q) * nf + j] = val

        fac_d = P(1) / (P(0) - CI * om)
        fac_14 = (P(3) * P(5) + CI * om * P(6)) / (om ** 2 - P(3) + CI * om * P(4))
        fac_13 = (CI / om) * (fac_14 + P(5))
        fac_16 = (P(8) * P(10) + CI * om * P(11)) / (om ** 2 - P(8) + CI * om * P(9))
        fac_15 = (CI / om) * (fac_16 + P(10))
        for c in range(3):
            put(c, 1, (fac_d * enpar[c]).real)
            put(c, 3, (fac_14 * enpar[c]).real)
            put(c, 2, (fac_13 * enpar[c]).real)
            put(c, 5, (fac_16 * enpar[c]).real)
            put(c, 4, (fac_15 * enpar[c]).real)

    def userfsrc(self, case):
        """userfsrc of the .usr: advance the sheet currents, then srcfh(c) -= fjn(j,c,1)"""
        import ctypes as C
        j = self.graphindex
        nf = case.nxzfl
        comps = (0, 1, 2) if self.imode == 3 else ((0, 1) if self.imode == 1 else (2,))

        def cb(tt, srcfhx, srcfhy, srcfhz, srcfex, srcfey, srcfez):
            case.L.ora_cem_graphene_current(C.byref(case.s), O.dp(self.fjn), O.dp(self.kfjn),
                                            O.dp(self.resfjn), O.dp(self.graphparams),
                                            O.dp(case.yconduc), O.ip(self.graphindex),
                                            int(j.size))
            src = (srcfhx, srcfhy, srcfhz)
            for c in comps:
                src[c][j] = src[c][j] - self.fjn[c * nf + j]

        return cb


def _case_graphene(imode, nx1, nel, box, param, dt):
    ldim = 3 if imode == 3 else 2
    bcs = ("P  ", "P  ", "PML", "PML", "P  ", "P  ")[:2 * ldim]
    mesh = O.box_mesh(nel, (box,) * ldim, bcs)
    u = _Graphene(imode)
    c = O.RefCase(mesh, nx1, imode=None if ldim == 3 else imode, upwind=True,
                  usrdat2=u.usrdat2, uservp=u.uservp, param=param)
    c.user = u
    c.set_dt(dt)
    c.usersol = lambda case, tt: u.usersol(case, tt)
    u.userini(c)
    c.set_callback("userinc", u.userinc(c))
    c.set_callback("userfsrc", u.userfsrc(c))
    c.nsteps = 1000
    return c


def case_3dgraphene(nx1=9, nel=(4, 12, 4)):
    """tests/3dgraphene (.box 4x12x4 on [-1,1]^3 rescaled to 5x10x5, BC P,P,PML,PML,P,P; N=8;
    dt=5e-3; 1000 steps; PML thick 2, order 3, referr 1e-10).  Tolerances 5e-4 / 5e-3 on
    hx,hz,ex,ez and 1e-14 / 5e-12 on hy,ey (3dgraphene.usr userchk)."""
    c = _case_graphene(3, nx1, nel, (-1.0, 1.0), {77: 2, 78: 3.0, 79: 1e-10}, -0.005)
    c.tol = dict(l2=[5e-4, 1e-14, 5e-4, 5e-4, 1e-14, 5e-4],
                 linf=[5e-3, 5e-12, 5e-3, 5e-3, 5e-12, 5e-3])
    return c


def case_2dgraphene(imode=1, nx1=9, nel=(4, 32), dt=0.2):
    """tests/2dgraphene (.box 4x32 on [-1500,1500]^2 rescaled to 5x10, BC P,P,PML,PML; N=8;
    param(12)=+0.2 -> CFL dt; 1000 steps; PML thick 6, order 3, referr 1e-15).  TE: 1e-7 / 5e-6 on
    hz, ex and 1e-14 / 5e-13 on ey; TM: the same on hx, ez / hy (2dgraphene.usr userchk)."""
    # NB for other meshes: the sheet ODEs are stiff (|lambda| ~ 680), so dt must stay below
    # ~ 5e-3 whatever the CFL number of the mesh says; the shipped mesh gives 5.04e-3
    c = _case_graphene(imode, nx1, nel, (-1500.0, 1500.0), {77: 6, 78: 3.0, 79: 1e-15}, dt)
    if imode == 1:
        c.tol = dict(l2=[0, 0, 1e-7, 1e-7, 1e-14, 0], linf=[0, 0, 5e-6, 5e-6, 5e-13, 0])
    else:
        c.tol = dict(l2=[1e-7, 1e-14, 0, 0, 0, 1e-7], linf=[5e-6, 5e-13, 0, 0, 0, 5e-6])
    return c


# ------------------------------------------------------------------------------------
# rotated elements: the same physical mesh with every element's local (r,s,t) frame turned by
# one of the 24 proper rotations -- neighbouring face lattices then run in different directions,
# which is what unstructured meshes (tests/cylwave, graphene) look like to the face pairing
# (SURVEY.md 8f rank 3).  Pure relabelling: the physical solution must not change.
# ------------------------------------------------------------------------------------
_PRE2SYM = (0, 1, 3, 2, 4, 5, 7, 6)  # preprocessor corner p -> symmetric index i+2j+4k
_FACE_SLOT = {(0, 0): 3, (0, 1): 1, (1, 0): 0, (1, 1): 2, (2, 0): 4, (2, 1): 5}  # (axis, side)


def proper_rotations():
    """the 24 signed axis permutations with determinant +1: list of (perm, sign), meaning the
    new local axis a runs along old axis perm[a], reversed if sign[a] < 0"""
    import itertools
    out = []
    for perm in itertools.permutations(range(3)):
        for sign in itertools.product((1, -1), repeat=3):
            m = np.zeros((3, 3))
            for a in range(3):
                m[a, perm[a]] = sign[a]
            if round(np.linalg.det(m)) == 1:
                out.append((perm, sign))
    return out


def rotate_elements(mesh, rot_of_elem):
    """new Mesh whose element e uses the local frame rot_of_elem[e] = (perm, sign)"""
    assert mesh.ldim == 3
    xc, yc, zc = mesh.xc.copy(), mesh.yc.copy(), mesh.zc.copy()
    vertex = mesh.vertex.copy()
    cbc = [list(r) for r in mesh.cbc]
    for e, (perm, sign) in enumerate(rot_of_elem):
        old_sym = {}
        for p in range(8):
            old_sym[_PRE2SYM[p]] = (mesh.xc[e, p], mesh.yc[e, p], mesh.zc[e, p])
        for p in range(8):
            ls = _PRE2SYM[p]
            new = (ls & 1, (ls >> 1) & 1, (ls >> 2) & 1)
            old = [0, 0, 0]
            for a in range(3):
                old[perm[a]] = new[a] if sign[a] > 0 else 1 - new[a]
            lo = old[0] + 2 * old[1] + 4 * old[2]
            xc[e, p], yc[e, p], zc[e, p] = old_sym[lo]
            vertex[e, ls] = mesh.vertex[e, lo]
        for a in range(3):
            for side in (0, 1):
                oside = side if sign[a] > 0 else 1 - side
                cbc[e][_FACE_SLOT[(a, side)]] = mesh.cbc[e][_FACE_SLOT[(perm[a], oside)]]
    return O.Mesh(3, xc, yc, zc, cbc, vertex)


def rotated_node_map(nx1, nelt, rot_of_elem):
    """index array m with new_field[m] == old_field: m[old flat node] = new flat node"""
    n = nx1
    m = np.zeros(n ** 3 * nelt, dtype=np.int64)
    idx = np.indices((n, n, n))  # idx[a][i',j',k'] new indices, array order [i'][j'][k']
    for e, (perm, sign) in enumerate(rot_of_elem):
        old = [None, None, None]
        for a in range(3):
            old[perm[a]] = idx[a] if sign[a] > 0 else (n - 1 - idx[a])
        new_flat = idx[0] + n * idx[1] + n * n * idx[2]
        old_flat = old[0] + n * old[1] + n * n * old[2]
        m[e * n ** 3 + old_flat.ravel()] = e * n ** 3 + new_flat.ravel()
    return m


def case_boxper_rotated(nel=(3, 3, 3), nx1=5, dt=-1e-3, seed=7):
    """case_boxper with every element's local frame turned by a (deterministic) pseudo-random
    proper rotation.  Returns (case, rot_of_elem)."""
    pi2 = 2 * 4.0 * math.atan(1.0)
    base = O.box_mesh(nel, ((0.0, pi2),) * 3, ("P  ",) * 6)
    rots = proper_rotations()
    rot_of_elem = [rots[(5 * e + seed) % 24] for e in range(base.nelt)]  # all 24 occur
    mesh = rotate_elements(base, rot_of_elem)
    c = O.RefCase(mesh, nx1, upwind=True)
    c.set_dt(dt)
    c.usersol = usersol_3dboxper
    shn, sen = usersol_3dboxper(c, 0.0)
    c.hn[:] = shn; c.en[:] = sen
    return c, rot_of_elem
