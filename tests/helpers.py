"""Shared test helpers: feed an oracle RefCase (the stand-in for the Fortran COMMON blocks)
through the C ABI and compare fields."""
import numpy as np

from nekcem_b200 import MaxwellB200
from nekcem_b200.api import GEOMETRY_ARRAYS


def arrays_from_refcase(c, elems=None):
    """COMMON-block arrays of a RefCase as the dict MaxwellB200.cem_maxwell_init expects.
    ``elems`` (sorted global element ids) restricts to one rank's partition."""
    d = {}
    if elems is None:
        elems = np.arange(c.nelt)
    elems = np.asarray(elems)
    vol = (elems[:, None] * c.nxyz + np.arange(c.nxyz)[None, :]).reshape(-1)
    nfp = c.nxzf * c.nfaces
    fac = (elems[:, None] * nfp + np.arange(nfp)[None, :]).reshape(-1)
    for name in GEOMETRY_ARRAYS:
        a = getattr(c, name)
        if name in ("dxm1", "w3mn"):
            d[name] = a.copy()
        elif a.size == c.npts:
            d[name] = a[vol].copy()
        else:
            d[name] = a[fac].copy()
    for name in ("permittivity", "permeability"):
        d[name] = getattr(c, name)[vol].copy()
    for name in ("hn", "en", "khn", "ken", "pmlsigma", "pmlbn", "pmldn"):
        a = getattr(c, name).reshape(3, c.npts)
        d[name] = a[:, vol].copy().reshape(-1)
    d["glo_num"] = c.glo_num[fac].copy()
    # cempec / pmlptr renumbered to the local element order
    loc = -np.ones(c.nelt, dtype=np.int64)
    loc[elems] = np.arange(elems.size)
    pec = c.cempec[:c.ncempec].astype(np.int64)
    pe = pec // nfp
    keep = loc[pe] >= 0
    d["cempec"] = loc[pe[keep]] * nfp + pec[keep] % nfp
    pml = c.pmlptr[:c.maxpml].astype(np.int64)
    d["pmlptr"] = loc[pml][loc[pml] >= 0]
    d["volvm1"] = c.volvm1
    return d


def incident_3ddielectric(c, elems=None):
    """The 3ddielectric `userinc` (tests/3ddielectric/3ddielectric.usr:6-50) as the arguments of
    MaxwellB200.set_incident: (face points, amp[6, ninc], phase, omega).  ``elems`` restricts
    and renumbers to one rank's partition."""
    u = c.user
    j = u.incindex
    k = c.cemface[j]
    eps = c.permittivity[k]; mu = c.permeability[k]
    eta = np.sqrt(mu / eps)
    ky = u.omega * np.sqrt(mu * eps)
    amp = np.zeros((6, j.size))
    amp[0] = -1.0 / eta   # incfhx -= uinc/eta
    amp[2] = 1.0          # incfhz += uinc
    amp[3] = eta          # incfex += eta*uinc
    amp[5] = 1.0          # incfez += uinc
    phase = -ky * c.ym1[k]
    if elems is not None:
        nfp = c.nxzf * c.nfaces
        loc = -np.ones(c.nelt, dtype=np.int64)
        loc[np.asarray(elems)] = np.arange(len(elems))
        keep = loc[j // nfp] >= 0
        j = loc[j[keep] // nfp] * nfp + j[keep] % nfp
        amp = amp[:, keep]; phase = phase[keep]
    return j, amp, phase, u.omega


def solver_from_refcase(c, device=0, incident=None, ade=None, strict=False):
    """ade = (kind, jn, kjn, params, index0) registers a Drude/Lorentz ADE; strict = the no-FMA
    instantiations (desc.strict)."""
    s = MaxwellB200(c.ldim, c.nx1, c.nelt, imode=c.imode, upwind=bool(c.s.ifupwind),
                    ifpec=c.ifpec, ifpml=c.ifpml, device=device, strict=strict)
    s.cem_maxwell_init(arrays_from_refcase(c))
    if incident is not None:
        s.set_incident(*incident)
    if ade is not None:
        kind, jn, kjn, params, index0 = ade
        (s.cem_maxwell_drude if kind == "drude" else s.cem_maxwell_lorentz)(jn, kjn, params, index0)
    s.setup()
    s.set_time(c.s.time, c.s.dt)
    return s


def rel_l2(a, b):
    """relative L2 difference of two field vectors (all components together)."""
    a = np.asarray(a); b = np.asarray(b)
    den = np.sqrt(np.sum(b * b))
    return float(np.sqrt(np.sum((a - b) ** 2)) / (den if den > 0 else 1.0))


def restrict_to_elems(c, elems, facepts=None, nodes=None):
    """Renumber global face points / nodes of RefCase ``c`` to the local numbering of the
    partition ``elems`` (sorted global element ids).  Returns (keep_mask, local_ids)."""
    elems = np.asarray(elems)
    loc = -np.ones(c.nelt, dtype=np.int64)
    loc[elems] = np.arange(elems.size)
    if facepts is not None:
        nfp = c.nxzf * c.nfaces
        j = np.asarray(facepts, dtype=np.int64)
        keep = loc[j // nfp] >= 0
        return keep, loc[j[keep] // nfp] * nfp + j[keep] % nfp
    j = np.asarray(nodes, dtype=np.int64)
    keep = loc[j // c.nxyz] >= 0
    return keep, loc[j[keep] // c.nxyz] * c.nxyz + j[keep] % c.nxyz


def planewave_args(c):
    """Arguments of MaxwellB200.cem_error_planewave for the layered-media cases of
    oracle/cases.py (3ddielectric, 2ddielectric, drude, lorentz, 3dgraphene, 2dgraphene): the
    `usersol` of their .usr files restated as two half-space plane waves with complex amplitudes
    and wavenumbers plus the graded PML decay.  Region 0 = upper half space, 1 = lower."""
    import math
    u = c.user
    kind = type(u).__name__
    region = (~u.upper).astype(np.uint8)
    inpml = (c.pmltag != 0).astype(np.uint8)
    order, referr = c.pmlorder, c.pmlreferr
    amp = np.zeros((2, 6), dtype=np.complex128)
    if kind == "_Dispersive":
        eta = [u.eta1, u.eta2]
        k = [u.k1, -u.k2]
        amp[0, 2], amp[0, 3] = u.refl, -u.eta1 * u.refl
        amp[1, 2], amp[1, 3] = u.tran, u.eta2 * u.tran
        eta_pml = [u.eta1, 0.0]
    else:
        e1, e2 = math.sqrt(u.mu1 / u.eps1), math.sqrt(u.mu2 / u.eps2)
        eta = [e1, e2]
        k = [u.omega * math.sqrt(u.mu1 * u.eps1), -u.omega * math.sqrt(u.mu2 * u.eps2)]
        imode = getattr(u, "imode", 3)
        if kind == "_Dielectric2D":
            te = (u.refl, u.tran) if imode == 1 else None
            tm = (u.refl, u.tran) if imode == 2 else None
        else:
            te = (u.reflte, u.trante) if imode in (3, 1) else None
            tm = (u.refltm, u.trantm) if imode in (3, 2) else None
        if te:
            amp[0, 2], amp[0, 3] = te[0], -e1 * te[0]
            amp[1, 2], amp[1, 3] = te[1], e2 * te[1]
        if tm:
            amp[0, 5], amp[0, 0] = tm[0], tm[0] / e1
            amp[1, 5], amp[1, 0] = tm[1], -tm[1] / e2
        eta_pml = eta
    d_u = c.pmlouter[3] - c.pmlinner[3]
    d_l = c.pmlinner[2] - c.pmlouter[2]
    pml = dict(order=order, eta=[0.0, 0.0], smax=[0.0, 0.0], d=[1.0, 1.0], y0=[0.0, 0.0],
               sign=[1.0, -1.0])
    for r, (d, y0) in enumerate(((d_u, c.pmlinner[3]), (d_l, c.pmlinner[2]))):
        if d > 0 and eta_pml[r] != 0.0:
            pml["eta"][r] = eta_pml[r]
            pml["smax"][r] = -(order + 1) * math.log(referr) / (2 * eta_pml[r] * d)
            pml["d"][r] = d
            pml["y0"][r] = y0
    return dict(omega=u.omega, k=k, amp=amp, region=region, inpml=inpml, pml=pml)


def planewave_numpy(c, args, tt):
    """the same formula in numpy (what the device kernel evaluates), for CPU checks of
    planewave_args against the cases' own usersol"""
    n = c.npts
    r = np.repeat(args["region"].astype(int), c.nxyz)
    pm = np.repeat(args["inpml"] != 0, c.nxyz)
    y = c.ym1
    p = args["pml"]
    g = lambda key: np.asarray(p[key], dtype=float)[r]
    with np.errstate(invalid="ignore"):
        fac = (g("smax") * g("d") / (p["order"] + 1)) * (g("sign") * (y - g("y0")) / g("d")) ** (p["order"] + 1)
    fac = np.where(pm, fac, 0.0)
    k = np.asarray(args["k"], dtype=complex)[r]
    uu = np.exp(1j * (k * y - args["omega"] * tt) - g("eta") * fac)
    out = [(args["amp"][r, cc] * uu).real for cc in range(6)]
    return np.concatenate(out[:3]), np.concatenate(out[3:])
