"""Driver for oracle/_ref/libnekcem_ref.so: the reference's own hot path (its Fortran
translated mechanically by oracle/f2c_lite.py + its own src/jl gather-scatter library,
recipe oracle/build_ref.py) run on the inputs of an oracle case.  TEST INFRASTRUCTURE ONLY.

What is the reference here and what is not: everything from `cem_maxwell_op_rk` down
(rk_c, cem_maxwell_op, cem_maxwell, maxwell_wght_curl, local_grad3/2, mxm -> mxfK,
restrict_to_face, flux2d/3d, gs_op_fields, flux_pec, add_flux_to_res, pml_step, invqmass,
rk_maxwell_ab, rk4_upd, rk_storage, cem_maxwell_drude/lorentz, cem_set_fc_ptr) executes the
reference's statements.  The COMMON-block inputs (geometry, masses, impedances, index lists,
initial fields) are filled from the oracle's setup, exactly as the drop-in library receives
them from the Fortran COMMONs; the .usr callbacks are supplied by the test.
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

from . import build_ref

c_dp = C.POINTER(C.c_double)
USERCB = C.CFUNCTYPE(None, c_dp, c_dp, c_dp, c_dp, c_dp, c_dp, c_dp)

_libs = {}


def available(kind: str = "base") -> bool:
    path = build_ref.build()
    if path is None:
        return False
    return kind == "base" or os.path.exists(build_ref.LIB_DROPIN)


def lib(kind: str = "base"):
    """kind 'base': the translated reference alone.  kind 'dropin': the same plus this repo's
    fixed-form shim fortran/cem_maxwell_b200_f77.F, linked against libnekcem_b200.so (its own
    copy of the COMMON blocks: the two libraries do not share state)."""
    if kind not in _libs:
        path = build_ref.build()
        if path is None:
            raise RuntimeError("oracle/_ref is not built and /root/reference is absent")
        if kind == "dropin":
            path = build_ref.LIB_DROPIN
        L = C.CDLL(path)
        L.ref_set_param.argtypes = [C.c_char_p, C.c_long]
        L.ref_sym.argtypes = [C.c_char_p, C.POINTER(C.c_int), C.POINTER(C.c_long)]
        L.ref_sym.restype = C.c_void_p
        L.ref_set_user.argtypes = [C.c_int, USERCB]
        L.ref_units.restype = C.c_char_p
        _libs[kind] = L
    return _libs[kind]


class ReferenceRun:
    """One process-wide instance at a time (the reference's state is global COMMON storage)."""

    def __init__(self, case, kind: str = "base", pad_elems: int = 0):
        """pad_elems > 0 dimensions the COMMON arrays for lelt = nelt + pad_elems elements, as a
        real SIZE file does (lelt = lelg/lpmin + 3, tests/3dboxper/SIZE:15): vector fields then
        have the leading dimension lpts1 > npts, which put() / field() honour."""
        L = lib(kind)
        self.L, self.case = L, case
        self.pad_elems = pad_elems
        ldim, nx1, nelt = case.ldim, case.nx1, case.nelt
        # SIZE: parameter (ldim, lxi, lelg ...) of this case; lelt = nelt exactly so that the
        # (lpts1,3) arrays have the oracle's layout (the reference pads lelt by 3 unused
        # elements, tests/3dboxper/SIZE:15)
        lelt = nelt + pad_elems
        for k, v in (("ldim", ldim), ("lxi", nx1 - 1), ("lelg", nelt), ("lpmin", 1),
                     ("lelv", lelt), ("lelt", lelt), ("lpts10", case.nxyz * lelt),
                     ("lxzfl10", case.nxzf * case.nfaces * lelt)):
            L.ref_set_param(k.encode(), v)
        L.ref_alloc()
        npts, nxzfl = case.npts, case.nxzfl
        self.lpts1 = int(self.get("lpts1"))
        assert self.lpts1 == case.nxyz * lelt and self.get("lxzfl1") == case.nxzf * case.nfaces * lelt
        nz1 = nx1 if ldim == 3 else 1
        # /DIMN/ (src/cem_drive.F:697-707)
        for k, v in (("nelt", nelt), ("nelv", nelt), ("nx1", nx1), ("ny1", nx1), ("nz1", nz1),
                     ("ndim", ldim), ("nxyz", case.nxyz), ("npts", npts), ("nxzf", case.nxzf),
                     ("nxzfl", nxzfl), ("nfaces", case.nfaces), ("imode", case.imode),
                     ("ifupwind", int(case.s.ifupwind)), ("ifcentral", int(case.s.ifcentral)),
                     ("ifpml", int(case.ifpml)), ("ifpec", int(case.ifpec)), ("ifrk45", 1),
                     ("ifrk22", 0), ("iffilter", 0), ("ifdealias", 0),
                     ("ifte", int(case.imode == 1)), ("iftm", int(case.imode == 2)),
                     ("ncemface", nxzfl), ("ncempec", case.ncempec), ("maxpml", case.maxpml),
                     ("istep", 0)):
            try:
                self.set(k, v)
            except KeyError:
                pass  # a COMMON scalar none of the translated routines reads
        self.set("dt", case.s.dt)
        self.set("time", case.s.time)
        try:
            self.set("pi", 4.0 * np.arctan(1.0))  # src/cem_drive.F: pi = 4.*atan(1.)
        except KeyError:
            pass
        # geometry, masses, impedances, fields: same names as the COMMON blocks
        for name in ("dxm1", "dxtm1", "w3mn", "rxmn", "rymn", "rzmn", "sxmn", "symn", "szmn",
                     "txmn", "tymn", "tzmn", "bmn", "unxm", "unym", "unzm", "aream", "hbm1",
                     "ebm1", "hn", "en", "khn", "ken", "permittivity", "permeability",
                     "pmlsigma", "pmlbn", "pmldn", "kpmlbn", "kpmldn"):
            try:
                self.put(name, getattr(case, name))
            except KeyError:
                pass
        for name in ("Y_0", "Y_1", "Z_0", "Z_1"):
            self.put(name.lower(), getattr(case, name))
        self.put("bm1", case.bmn)
        # index lists are 1-based in the reference
        self.put("cemface", case.cemface + 1)
        if case.ncempec:
            self.put("cempec", case.cempec[:case.ncempec] + 1)
        if case.maxpml:
            self.put("pmlptr", case.pmlptr[:case.maxpml] + 1)
        # gs_setup(gsh_face, glo_num, nxzfl, comm, np)  (src/nek5_connect11.F:2217-2223)
        h = C.c_int(-1)
        n = C.c_int(nxzfl)
        comm, np_ = C.c_int(0), C.c_int(1)
        ids = np.ascontiguousarray(case.glo_num, dtype=np.int64)
        L.gs_setup_(C.byref(h), ids.ctypes.data_as(C.c_void_p), C.byref(n), C.byref(comm),
                    C.byref(np_))
        self.gsh = h.value
        self.set("gsh_face", self.gsh)
        # what the drop-in shim reads besides the arrays above: the face ids still resident in
        # COMMON /c_is1/ glo_num (src/nek5_connect11.F:31), and the rank / size
        self.put_opt("glo_num", ids)
        for k, v in (("nid", 0), ("np", 1)):
            try:
                self.set(k, v)
            except KeyError:
                pass
        L.rk_storage_()
        # setup_topo's face tables (src/nek5_connect11.F:1046-1093, 1440-1528).  dsset keeps
        # the dimensions of its last call in SAVEd variables and returns early when they
        # repeat; ref_alloc has just re-created the COMMON arrays, so force a refresh.
        L.initds_()
        one = C.c_int(1)
        L.dsset_(C.byref(one), C.byref(one), C.byref(one))
        L.dsset_(C.byref(C.c_int(nx1)), C.byref(C.c_int(nx1)), C.byref(C.c_int(nz1)))
        self._cbs = {}
        self.istep = 0

    # ---- COMMON access ------------------------------------------------------------------
    def _sym_kind(self, name):
        isint, cnt = C.c_int(), C.c_long()
        p = self.L.ref_sym(name.encode(), C.byref(isint), C.byref(cnt))
        if not p:
            raise KeyError(name)
        return p, isint.value, cnt.value

    def _sym(self, name):
        p, kind, cnt = self._sym_kind(name)
        return p, kind == 1, cnt

    def view(self, name):
        p, kind, cnt = self._sym_kind(name)
        isint = kind == 1
        if kind == 3:
            return np.ctypeslib.as_array(C.cast(p, C.POINTER(C.c_longlong)), shape=(cnt,))
        if kind == 4:   # complex (double _Complex)
            return np.ctypeslib.as_array(C.cast(p, C.POINTER(C.c_double)),
                                         shape=(2 * cnt,)).view(np.complex128)
        t = C.c_int if isint else C.c_double
        return np.ctypeslib.as_array(C.cast(p, C.POINTER(t)), shape=(cnt,))

    def get(self, name):
        return self.view(name)[0]

    def set(self, name, v):
        self.view(name)[0] = v

    def put(self, name, arr):
        v = self.view(name)
        a = np.asarray(arr).reshape(-1)
        assert a.size <= v.size, (name, a.size, v.size)
        npts = self.case.npts
        if self.pad_elems and a.size == 3 * npts and v.size == 3 * self.lpts1:
            for k in range(3):           # (lpts1,3) Fortran array <- compact (npts,3)
                v[k * self.lpts1:k * self.lpts1 + npts] = a[k * npts:(k + 1) * npts]
            return
        v[:a.size] = a

    def field(self, name):
        """compact (npts,3) copy of an (lpts1,3) COMMON array"""
        v, npts = self.view(name), self.case.npts
        return np.concatenate([v[k * self.lpts1:k * self.lpts1 + npts] for k in range(3)])

    def view_char(self, name):
        p, kind, cnt = self._sym_kind(name)
        assert kind == 2, name
        return np.ctypeslib.as_array(C.cast(p, C.POINTER(C.c_uint8)), shape=(cnt,))

    def set_cbc(self, cbc):
        """CBC(6,LELT,0:LDIMT1) character*3 (src/INPUT:147-150), field 1 = the E&M boundary
        conditions; cbc[e][f] are the 3-character flags in preprocessor face order"""
        v = self.view_char("cbc")
        lelt = self.case.nelt
        v[:] = ord(" ")
        for e, row in enumerate(cbc):
            for f, s in enumerate(row):
                off = (f + 6 * (e + lelt * 1)) * 3
                v[off:off + 3] = np.frombuffer(s.ljust(3)[:3].encode(), dtype=np.uint8)

    def put_opt(self, name, arr):
        """put() for a COMMON array that the translated routines may not reference"""
        try:
            self.put(name, arr)
        except KeyError:
            pass

    # ---- callbacks ------------------------------------------------------------------------
    def set_callback(self, which: str, fn):
        """fn(tt, a1..a6): numpy views, argument order of the reference's call site"""
        n = self.case.nxzfl if which in ("userinc", "userfsrc") else self.case.npts

        def tramp(t, a1, a2, a3, a4, a5, a6):
            arrs = [np.ctypeslib.as_array(a, shape=(n,)) for a in (a1, a2, a3, a4, a5, a6)]
            fn(t[0], *arrs)

        cb = USERCB(tramp)
        self._cbs[which] = cb
        self.L.ref_set_user({"userinc": 0, "usersrc": 1, "userfsrc": 2}[which], cb)

    def clear_callbacks(self):
        for k in (0, 1, 2):
            self.L.ref_set_user(k, C.cast(None, USERCB))
        self._cbs = {}

    # ---- stepping -------------------------------------------------------------------------
    def step(self, nsteps: int = 1):
        """the body of the time loop, src/cem_drive.F:622-640"""
        for _ in range(nsteps):
            self.istep += 1
            try:
                self.set("istep", self.istep)
            except KeyError:
                pass
            self.L.cem_maxwell_op_rk_()
            self.set("time", self.get("time") + self.get("dt"))

    def stage(self, rkstep: int):
        """one pass of the loop in cem_maxwell_op_rk (src/cem_maxwell.F:337-341)"""
        i = C.c_int(rkstep)
        self.set("rkstep", rkstep)
        self.L.rk_c_(C.byref(i))
        self.L.cem_maxwell_op_()
        self.L.rk_maxwell_ab_(C.byref(i))

    @property
    def hn(self):
        return self.view("hn")[:3 * self.case.npts]

    @property
    def en(self):
        return self.view("en")[:3 * self.case.npts]

    def close(self):
        self.clear_callbacks()
        h = C.c_int(self.gsh)
        self.L.gs_free_(C.byref(h))
