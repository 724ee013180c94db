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


def solver_from_refcase(c, device=0, incident=None, ade=None):
    """ade = (kind, jn, kjn, params, index0) registers a Drude/Lorentz ADE."""
    s = MaxwellB200(c.ldim, c.nx1, c.nelt, imode=c.imode, upwind=bool(c.s.ifupwind),
                    ifpec=c.ifpec, ifpml=c.ifpml, device=device)
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
