"""Host-side mirror of the reference's Maxwell operator interface over libnekcem_b200.so.

The reference's "operator API" for this path is a set of Fortran subroutine names plus
COMMON blocks.  This module is the Python stand-in for that Fortran host
side, used by the tests, ``bench.py`` and ``__graft_entry__``: same names, same argument
meaning, same error behaviour (an error in the library raises ``NekcemB200Error`` where
the reference would ``call exitt(1)``).

    cem_maxwell_init   -> MaxwellB200.cem_maxwell_init(arrays)   (src/cem_maxwell.F:65-191)
    acc_copy_all_in    -> done by cem_maxwell_init / set_array   (src/cem_drive.F:397-480)
    cem_maxwell_op_rk  -> MaxwellB200.cem_maxwell_op_rk(nsteps)  (src/cem_maxwell.F:327-345)
    !$ACC UPDATE HOST  -> MaxwellB200.hn / .en / get_array       (tests/3dboxper/3dboxper.usr:199)
    cem_error          -> MaxwellB200.cem_error                  (src/cem_common.F:1335-1355)

There is no CPU fallback: the CUDA library must be present and, for compute, a B200.
Index arrays are 0-based on the Python side and converted to the ABI's 1-based form here.
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIBPATH = os.environ.get("NEKCEM_B200_LIB") or os.path.join(_HERE, "lib", "libnekcem_b200.so")
ABI_VERSION = 1

ARRAY_IDS = {name: i for i, name in enumerate([
    "dxm1", "w3mn", "rxmn", "rymn", "rzmn", "sxmn", "symn", "szmn", "txmn", "tymn", "tzmn",
    "bmn", "hbm1", "ebm1", "unxm", "unym", "unzm", "aream", "Y_0", "Y_1", "Z_0", "Z_1",
    "hn", "en", "khn", "ken", "permittivity", "permeability", "pmlsigma", "pmlbn", "pmldn",
    "kpmlbn", "kpmldn", "xmn", "ymn", "zmn", "yconduc"])}

GEOMETRY_ARRAYS = ["dxm1", "w3mn", "rxmn", "rymn", "rzmn", "sxmn", "symn", "szmn", "txmn",
                   "tymn", "tzmn", "bmn", "hbm1", "ebm1", "unxm", "unym", "unzm", "aream",
                   "Y_0", "Y_1", "Z_0", "Z_1"]
PML_ARRAYS = ["permittivity", "permeability", "pmlsigma", "pmlbn", "pmldn"]


class NekcemB200Error(RuntimeError):
    pass


class PlaneWave(C.Structure):
    """nekcem_b200_planewave (include/nekcem_b200.h)"""
    _fields_ = [("omega", C.c_double), ("k_re", C.c_double * 2), ("k_im", C.c_double * 2),
                ("amp_re", (C.c_double * 6) * 2), ("amp_im", (C.c_double * 6) * 2),
                ("pml_eta", C.c_double * 2), ("pml_smax", C.c_double * 2),
                ("pml_d", C.c_double * 2), ("pml_y0", C.c_double * 2),
                ("pml_sign", C.c_double * 2), ("pml_order", C.c_double)]


class Desc(C.Structure):
    _fields_ = [(n, C.c_int32) for n in (
        "abi_version", "ldim", "nx1", "nelt", "imode", "ifupwind", "ifpec", "ifpml", "device",
        "strict", "rank", "nranks")]


_lib = None
c_dp = C.POINTER(C.c_double)
c_i64p = C.POINTER(C.c_int64)
c_i32p = C.POINTER(C.c_int32)


def lib():
    """Load libnekcem_b200.so (fails loudly if the CUDA extension has not been built)."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIBPATH):
            raise NekcemB200Error(
                f"{LIBPATH} is missing: build it with `python -c 'import __graft_entry__ as g; "
                "g.build()'` (nvcc, sm_100a).  There is no CPU fallback.")
        L = C.CDLL(LIBPATH)
        L.nekcem_b200_last_error.restype = C.c_char_p
        L.nekcem_b200_create.argtypes = [C.POINTER(Desc), C.POINTER(C.c_int)]
        L.nekcem_b200_destroy.argtypes = [C.c_int]
        L.nekcem_b200_set_array.argtypes = [C.c_int, C.c_int, c_dp, C.c_int64]
        L.nekcem_b200_get_array.argtypes = [C.c_int, C.c_int, c_dp, C.c_int64]
        L.nekcem_b200_set_faces.argtypes = [C.c_int, c_i64p, C.c_int64, c_i32p, C.c_int32]
        L.nekcem_b200_set_pml.argtypes = [C.c_int, c_i32p, C.c_int32]
        L.nekcem_b200_comm_unique_id.argtypes = [C.c_char_p]
        L.nekcem_b200_comm_init.argtypes = [C.c_int, C.c_char_p]
        L.nekcem_b200_face_singletons.argtypes = [C.c_int, c_i64p, C.c_int64, c_i64p]
        L.nekcem_b200_face_remote.argtypes = [C.c_int, c_i64p, c_i64p]
        L.nekcem_b200_plan_npeers.argtypes = [C.c_int, c_i32p, c_i64p]
        L.nekcem_b200_plan_peer.argtypes = [C.c_int, C.c_int32, c_i32p, c_i64p, c_i64p]
        L.nekcem_b200_plan_vmap.argtypes = [C.c_int, c_i32p, C.c_int64]
        L.nekcem_b200_plan_elements.argtypes = [C.c_int, c_i32p, c_i32p]
        L.nekcem_b200_setup.argtypes = [C.c_int]
        L.nekcem_b200_set_incident.argtypes = [C.c_int, C.c_int32, c_i32p, c_dp, c_dp, C.c_double]
        L.nekcem_b200_set_volume_source.argtypes = [C.c_int, C.c_int, c_dp, C.c_double,
                                                    C.c_double, C.c_double]
        L.nekcem_b200_set_drude.argtypes = [C.c_int, c_dp, c_dp, c_dp, c_i32p, C.c_int32]
        L.nekcem_b200_set_lorentz.argtypes = [C.c_int, c_dp, c_dp, c_dp, c_i32p, C.c_int32]
        L.nekcem_b200_get_ade.argtypes = [C.c_int, c_dp, c_dp]
        L.nekcem_b200_set_graphene.argtypes = [C.c_int, c_dp, c_dp, c_dp, c_dp, c_i32p, C.c_int32]
        L.nekcem_b200_get_graphene.argtypes = [C.c_int, c_dp, c_dp]
        L.nekcem_b200_set_option.argtypes = [C.c_int, C.c_char_p, C.c_int]
        L.nekcem_b200_error_sums_mode.argtypes = [C.c_int, C.POINTER(C.c_int32), c_dp, c_dp, c_dp,
                                                  c_dp, c_dp]
        L.nekcem_b200_error_sums_planewave.argtypes = [C.c_int, C.POINTER(PlaneWave),
                                                       C.POINTER(C.c_ubyte), C.POINTER(C.c_ubyte),
                                                       C.c_double, c_dp, c_dp]
        L.nekcem_b200_set_leading_dims.argtypes = [C.c_int, C.c_int64, C.c_int64]
        L.nekcem_b200_set_array_ld.argtypes = [C.c_int, C.c_int, c_dp, C.c_int64]
        L.nekcem_b200_get_array_ld.argtypes = [C.c_int, C.c_int, c_dp, C.c_int64]
        L.nekcem_b200_set_rk_coefficients.argtypes = [C.c_int, c_dp, c_dp, c_dp]
        L.nekcem_b200_get_rk_coefficients.argtypes = [C.c_int, c_dp, c_dp, c_dp]
        L.nekcem_b200_set_filter.argtypes = [C.c_int, c_dp]
        L.nekcem_b200_apply_filter.argtypes = [C.c_int]
        L.nekcem_b200_vtk_payload.argtypes = [C.c_int, C.c_int, C.c_int, C.c_void_p]
        L.nekcem_b200_restart_ingest.argtypes = [C.c_int, C.c_int, C.c_int, C.c_void_p]
        L.nekcem_b200_geometry_info.argtypes = [C.c_int, C.POINTER(C.c_int64), C.POINTER(C.c_int32)]
        L.nekcem_b200_set_time.argtypes = [C.c_int, C.c_double, C.c_double]
        L.nekcem_b200_get_time.argtypes = [C.c_int, c_dp]
        L.nekcem_b200_step.argtypes = [C.c_int, C.c_int]
        L.nekcem_b200_step_streamed.argtypes = [C.c_int, c_dp, c_dp, c_dp, c_dp]
        L.nekcem_b200_device_count.argtypes = []
        L.nekcem_b200_transport.argtypes = [C.c_int, C.POINTER(C.c_int32)]
        L.nekcem_b200_stage.argtypes = [C.c_int, C.c_int]
        L.nekcem_b200_synchronize.argtypes = [C.c_int]
        L.nekcem_b200_apply_rhs.argtypes = [C.c_int, C.c_double]
        L.nekcem_b200_stage_pack.argtypes = [C.c_int, C.c_int]
        L.nekcem_b200_stage_compute.argtypes = [C.c_int, C.c_int]
        L.nekcem_b200_halo_exchange_local.argtypes = [C.c_int, C.c_int]
        L.nekcem_b200_halo_buffers.argtypes = [C.c_int, C.c_int32, C.POINTER(C.c_void_p),
                                               C.POINTER(C.c_void_p), c_i64p]
        L.nekcem_b200_error_sums.argtypes = [C.c_int, c_dp, c_dp, c_dp, c_dp]
        L.nekcem_b200_last_step_ms.argtypes = [C.c_int, C.POINTER(C.c_float), c_i64p]
        L.nekcem_b200_algorithmic_bytes.argtypes = [C.c_int, c_dp]
        _lib = L
    return _lib


def _chk(rc):
    if rc != 0:
        raise NekcemB200Error(lib().nekcem_b200_last_error().decode())


def _dp(a):
    assert a.dtype == np.float64 and a.flags["C_CONTIGUOUS"], "need contiguous float64"
    return a.ctypes.data_as(c_dp)


def comm_unique_id() -> bytes:
    buf = C.create_string_buffer(128)
    _chk(lib().nekcem_b200_comm_unique_id(buf))
    return buf.raw


class MaxwellB200:
    """One context = one rank's share of the mesh on one GPU."""

    def __init__(self, ldim: int, nx1: int, nelt: int, imode: int = 3, upwind: bool = True,
                 ifpec: bool = False, ifpml: bool = False, device: int = 0, rank: int = 0,
                 nranks: int = 1, strict: bool = False):
        """strict: the no-FMA instantiations of the stage kernels (3D: the slab formulation at every
        order; 2D) and of the graphene kernel: every product and sum rounded separately, as in the
        reference's x86-64 build"""
        self.L = lib()
        d = Desc(ABI_VERSION, ldim, nx1, nelt, imode, int(upwind), int(ifpec), int(ifpml),
                 device, int(strict), rank, nranks)
        h = C.c_int(-1)
        _chk(self.L.nekcem_b200_create(C.byref(d), C.byref(h)))
        self.h = h.value
        self.ldim, self.nx1, self.nelt = ldim, nx1, nelt
        self.nxyz = nx1 ** 3 if ldim == 3 else nx1 ** 2
        self.nxzf = nx1 ** 2 if ldim == 3 else nx1
        self.nfaces = 2 * ldim
        self.npts = self.nxyz * nelt
        self.nxzfl = self.nxzf * self.nfaces * nelt
        self.rank, self.nranks = rank, nranks
        self.volvm1 = None

    # -- lifetime ----------------------------------------------------------------------
    def close(self):
        if getattr(self, "h", None) is not None and self.h >= 0:
            self.L.nekcem_b200_destroy(self.h)
            self.h = -1

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # -- acc_copy_all_in / UPDATE HOST ---------------------------------------------------
    def set_array(self, name: str, arr: np.ndarray):
        a = np.ascontiguousarray(arr, dtype=np.float64).reshape(-1)
        _chk(self.L.nekcem_b200_set_array(self.h, ARRAY_IDS[name], _dp(a), a.size))

    def get_array(self, name: str) -> np.ndarray:
        which = ARRAY_IDS[name]
        if name in ("dxm1",):
            n = self.nx1 * self.nx1
        elif name == "w3mn":
            n = self.nxyz
        elif name in ("unxm", "unym", "unzm", "aream", "Y_0", "Y_1", "Z_0", "Z_1", "yconduc"):
            n = self.nxzfl
        elif name in ("hn", "en", "khn", "ken", "pmlsigma", "pmlbn", "pmldn", "kpmlbn",
                      "kpmldn"):
            n = 3 * self.npts
        else:
            n = self.npts
        out = np.zeros(n)
        _chk(self.L.nekcem_b200_get_array(self.h, which, _dp(out), n))
        return out

    @property
    def hn(self):
        return self.get_array("hn")

    @property
    def en(self):
        return self.get_array("en")

    def set_faces(self, glo_num: np.ndarray, cempec0: np.ndarray):
        """glo_num: face-point global ids (int64, nxzfl); cempec0: 0-based PEC/PML face points."""
        g = np.ascontiguousarray(glo_num, dtype=np.int64)
        p = np.ascontiguousarray(np.asarray(cempec0, dtype=np.int64) + 1, dtype=np.int32)
        _chk(self.L.nekcem_b200_set_faces(self.h, g.ctypes.data_as(c_i64p), g.size,
                                          p.ctypes.data_as(c_i32p), p.size))

    def set_pml(self, pmlptr0: np.ndarray):
        p = np.ascontiguousarray(np.asarray(pmlptr0, dtype=np.int64) + 1, dtype=np.int32)
        _chk(self.L.nekcem_b200_set_pml(self.h, p.ctypes.data_as(c_i32p), p.size))

    # -- multi-rank plumbing ---------------------------------------------------------------
    def comm_init(self, uid: bytes):
        _chk(self.L.nekcem_b200_comm_init(self.h, uid))

    def face_singletons(self) -> np.ndarray:
        cnt = C.c_int64(0)
        _chk(self.L.nekcem_b200_face_singletons(self.h, None, 0, C.byref(cnt)))
        ids = np.zeros(max(cnt.value, 1), dtype=np.int64)
        _chk(self.L.nekcem_b200_face_singletons(self.h, ids.ctypes.data_as(c_i64p), ids.size,
                                                C.byref(cnt)))
        return ids[:cnt.value]

    def face_remote(self, counts, all_ids):
        c = np.ascontiguousarray(counts, dtype=np.int64)
        a = np.ascontiguousarray(all_ids, dtype=np.int64)
        if a.size == 0:
            a = np.zeros(1, dtype=np.int64)
        _chk(self.L.nekcem_b200_face_remote(self.h, c.ctypes.data_as(c_i64p),
                                            a.ctypes.data_as(c_i64p)))

    def plan(self):
        """Exchange plan: (vmapP, [(peer_rank, send_facepts)], nhalo, n_interior, n_boundary)."""
        npeers = C.c_int32(0); nhalo = C.c_int64(0)
        _chk(self.L.nekcem_b200_plan_npeers(self.h, C.byref(npeers), C.byref(nhalo)))
        peers = []
        for ipeer in range(npeers.value):
            r = C.c_int32(0); cnt = C.c_int64(0)
            _chk(self.L.nekcem_b200_plan_peer(self.h, ipeer, C.byref(r), C.byref(cnt), None))
            fp = np.zeros(cnt.value, dtype=np.int64)
            _chk(self.L.nekcem_b200_plan_peer(self.h, ipeer, C.byref(r), C.byref(cnt),
                                              fp.ctypes.data_as(c_i64p)))
            peers.append((r.value, fp))
        vm = np.zeros(self.nxzfl, dtype=np.int32)
        _chk(self.L.nekcem_b200_plan_vmap(self.h, vm.ctypes.data_as(c_i32p), vm.size))
        ni = C.c_int32(0); nb = C.c_int32(0)
        _chk(self.L.nekcem_b200_plan_elements(self.h, C.byref(ni), C.byref(nb)))
        return vm, peers, nhalo.value, ni.value, nb.value

    # -- cem_maxwell_init ----------------------------------------------------------------
    def cem_maxwell_init(self, arrays: dict, free_after_upload: bool = False):
        """Upload the COMMON arrays the hot path reads (what acc_copy_all_in lists,
        src/cem_drive.F:430-458) and the face connectivity.  ``arrays`` maps the reference's
        array names to numpy arrays; 'glo_num' (int64) and 'cempec' (0-based) describe the
        faces; 'pmlptr' (0-based) the PML elements."""
        for name in GEOMETRY_ARRAYS:
            self.set_array(name, arrays[name])
            if free_after_upload:
                arrays[name] = None
        if len(arrays.get("pmlptr", ())) > 0:
            for name in PML_ARRAYS:
                if arrays.get(name) is not None:
                    self.set_array(name, arrays[name])
            self.set_pml(arrays["pmlptr"])
        for name in ("hn", "en", "khn", "ken"):
            a = arrays.get(name)
            if a is not None:
                self.set_array(name, a)
        if free_after_upload:
            for name in ("hn", "en"):
                if name in arrays:
                    arrays[name] = None
        self.set_faces(arrays["glo_num"], arrays.get("cempec", np.zeros(0, dtype=np.int64)))
        if "volvm1" in arrays:
            self.volvm1 = float(arrays["volvm1"])

    def setup(self):
        _chk(self.L.nekcem_b200_setup(self.h))

    def set_incident(self, facepts0, amp, phase, omega):
        """userinc hook: on the 0-based face points ``facepts0`` the own trace gets
        ``amp[comp, q] * cos(phase[q] - omega*rktime)`` added in every stage (comp 0..2 = H,
        3..5 = E), as tests/3ddielectric/3ddielectric.usr:6-50 does.  Call before setup()."""
        fp = np.ascontiguousarray(np.asarray(facepts0, dtype=np.int64) + 1, dtype=np.int32)
        a = np.ascontiguousarray(amp, dtype=np.float64).reshape(-1)
        ph = np.ascontiguousarray(phase, dtype=np.float64).reshape(-1)
        assert a.size == 6 * fp.size and ph.size == fp.size
        _chk(self.L.nekcem_b200_set_incident(self.h, fp.size, fp.ctypes.data_as(c_i32p), _dp(a),
                                             _dp(ph), float(omega)))

    def set_volume_source(self, comp, profile, amp, omega, phase):
        p = None if profile is None else _dp(np.ascontiguousarray(profile, dtype=np.float64))
        _chk(self.L.nekcem_b200_set_volume_source(self.h, comp, p, amp, omega, phase))

    def _set_ade(self, fn, ncomp, npar, jn, kjn, params, index0):
        idx = np.ascontiguousarray(np.asarray(index0, dtype=np.int64) + 1, dtype=np.int32)
        par = np.ascontiguousarray(params, dtype=np.float64).reshape(-1)
        assert par.size == npar * self.npts
        ptr = []
        for a in (jn, kjn):
            if a is None:
                ptr.append(None)
            else:
                a = np.ascontiguousarray(a, dtype=np.float64).reshape(-1)
                assert a.size == ncomp * self.npts
                ptr.append(a)
        self._ade_ncomp = ncomp
        _chk(fn(self.h, None if ptr[0] is None else _dp(ptr[0]),
                None if ptr[1] is None else _dp(ptr[1]), _dp(par),
                idx.ctypes.data_as(c_i32p), idx.size))

    def cem_maxwell_drude(self, jn, kjn, params, dindex0):
        """Registers the Drude ADE that the reference's usersrc runs every stage through
        ``cem_maxwell_drude(jn,kjn,resjn,params,dindex,n)`` (src/cem_maxwell.F:3095-3147):
        jn,kjn (npts,3) or None, params (npts,2), dindex0 0-based nodes.  Call before setup()."""
        self._set_ade(self.L.nekcem_b200_set_drude, 3, 2, jn, kjn, params, dindex0)

    def cem_maxwell_lorentz(self, jn, kjn, params, lindex0):
        """``cem_maxwell_lorentz`` (src/cem_maxwell.F:3149-3211): jn,kjn (npts,3,2), params
        (npts,3)."""
        self._set_ade(self.L.nekcem_b200_set_lorentz, 6, 3, jn, kjn, params, lindex0)

    def get_ade(self):
        """(jn, kjn) of the registered ADE, downloaded from the device."""
        jn = np.zeros(self._ade_ncomp * self.npts); kjn = np.zeros_like(jn)
        _chk(self.L.nekcem_b200_get_ade(self.h, _dp(jn), _dp(kjn)))
        return jn, kjn

    def cem_graphene_current(self, fjn, kfjn, params, yconduc, gindex0):
        """Registers the graphene sheets that the reference's userfsrc advances every stage
        through ``cem_3d_graphene_current / cem_te_graphene_current / cem_tm_graphene_current
        (fjn,kfjn,resfjn,params,gindex,n)`` (src/cem_maxwell.F:2827-3093; the variant follows
        from imode) and whose total current it subtracts from its -(n x H) face source
        (tests/3dgraphene/3dgraphene.usr:236-266).  fjn,kfjn (nxzfl,3,6) or None, params
        (nxzfl,12), yconduc (nxzfl) = COMMON /EMWAVE/ yconduc (or None when the array "yconduc"
        was uploaded), gindex0 0-based face points.  Call before setup()."""
        idx = np.ascontiguousarray(np.asarray(gindex0, dtype=np.int64) + 1, dtype=np.int32)
        nf = self.nxzfl
        par = np.ascontiguousarray(params, dtype=np.float64).reshape(-1)
        if par.size != 12 * nf:
            raise NekcemB200Error(f"graphene params: {par.size} values, expected {12 * nf}")
        ptr = []
        for a in (fjn, kfjn):
            if a is None:
                ptr.append(None)
            else:
                a = np.ascontiguousarray(a, dtype=np.float64).reshape(-1)
                if a.size != 18 * nf:
                    raise NekcemB200Error(f"graphene state: {a.size} values, expected {18 * nf}")
                ptr.append(a)
        yc = None
        if yconduc is not None:
            yc = np.ascontiguousarray(yconduc, dtype=np.float64).reshape(-1)
            if yc.size != nf:
                raise NekcemB200Error(f"yconduc: {yc.size} values, expected {nf}")
        _chk(self.L.nekcem_b200_set_graphene(
            self.h, None if ptr[0] is None else _dp(ptr[0]),
            None if ptr[1] is None else _dp(ptr[1]), _dp(par), None if yc is None else _dp(yc),
            idx.ctypes.data_as(c_i32p), idx.size))

    def get_graphene(self):
        """(fjn, kfjn) as (nxzfl,3,6) flat arrays: the listed face points downloaded from the
        device, zeros elsewhere."""
        fjn = np.zeros(18 * self.nxzfl); kfjn = np.zeros_like(fjn)
        _chk(self.L.nekcem_b200_get_graphene(self.h, _dp(fjn), _dp(kfjn)))
        return fjn, kfjn

    def set_rk_coefficients(self, a, b, c):
        """COMMON /RKCOEF/ rk4a(5), rk4b(5), rk4c(6) as rk_storage left them (src/cem_common.F:78-114)"""
        a, b, c = (np.ascontiguousarray(x, dtype=np.float64) for x in (a, b, c))
        assert a.size == 5 and b.size == 5 and c.size == 6
        _chk(self.L.nekcem_b200_set_rk_coefficients(self.h, _dp(a), _dp(b), _dp(c)))

    def get_rk_coefficients(self):
        a, b, c = np.zeros(5), np.zeros(5), np.zeros(6)
        _chk(self.L.nekcem_b200_get_rk_coefficients(self.h, _dp(a), _dp(b), _dp(c)))
        return a, b, c

    def set_filter(self, intv):
        """param(18) = 1: every time step ends with q_filter (src/nek5_filter.F:2-144); intv is
        the nx1 x nx1 matrix of the reference's build_new_filter, column-major (None: off)"""
        if intv is None:
            _chk(self.L.nekcem_b200_set_filter(self.h, None))
            return
        f = np.ascontiguousarray(intv, dtype=np.float64).reshape(-1)
        if f.size != self.nx1 * self.nx1:
            raise NekcemB200Error(f"filter matrix: {f.size} values, expected {self.nx1 ** 2}")
        _chk(self.L.nekcem_b200_set_filter(self.h, _dp(f)))

    def apply_filter(self):
        _chk(self.L.nekcem_b200_apply_filter(self.h))

    def vtk_payload(self, which: str, as_double: bool = False) -> bytes:
        """The byte payload of the VTK "VECTORS" block cem_out writes for EN ('en') or HN ('hn'):
        per node three values, cast to float32 unless as_double, big-endian
        (vtk_nonswap_field + writefield4[_double], src/io_dumpvtk.F:858-878, src/io_co.c:414-536),
        assembled on the device."""
        w = {"en": 0, "hn": 1}[which]
        buf = np.empty(3 * self.npts * (8 if as_double else 4), dtype=np.uint8)
        _chk(self.L.nekcem_b200_vtk_payload(self.h, w, int(as_double), buf.ctypes.data_as(C.c_void_p)))
        return buf.tobytes()

    def restart_ingest(self, which: str, payload: bytes, as_double: bool = True):
        """The field part of ``restart_swap`` (src/io.F:637-781): fill EN ('en') or HN ('hn') from
        the big-endian "VECTORS" section of a restart file (float32 unless as_double), this rank's
        element order; byte swap, cast and de-interleave run on the device."""
        w = {"en": 0, "hn": 1}[which]
        buf = np.frombuffer(payload, dtype=np.uint8)
        assert buf.size == 3 * self.npts * (8 if as_double else 4)
        _chk(self.L.nekcem_b200_restart_ingest(self.h, w, int(as_double),
                                               buf.ctypes.data_as(C.c_void_p)))

    def set_option(self, name: str, value: int):
        _chk(self.L.nekcem_b200_set_option(self.h, name.encode(), int(value)))

    # -- time stepping -------------------------------------------------------------------
    def set_time(self, time: float, dt: float):
        _chk(self.L.nekcem_b200_set_time(self.h, time, dt))

    @property
    def time(self) -> float:
        t = C.c_double(0)
        _chk(self.L.nekcem_b200_get_time(self.h, C.byref(t)))
        return t.value

    def cem_maxwell_op_rk(self, nsteps: int = 1, sync: bool = True):
        """nsteps time steps of the 5-stage LSRK (src/cem_maxwell.F:327-345)."""
        _chk(self.L.nekcem_b200_step(self.h, nsteps))
        if sync:
            self.synchronize()

    step = cem_maxwell_op_rk

    def transport(self) -> str:
        """transport of the inter-rank face exchange chosen at setup"""
        k = C.c_int32(0)
        _chk(self.L.nekcem_b200_transport(self.h, C.byref(k)))
        return {0: "none", 1: "nccl-sendrecv", 2: "peer-memory-push"}[int(k.value)]

    def step_streamed(self, hn_in=None, en_in=None, hn_out=None, en_out=None):
        """One time step on a stream of host inputs (nekcem_b200_step_streamed): uploads
        (hn_in, en_in), advances the state the previous call uploaded, returns the result the
        previous call computed into (hn_out, en_out).  Arrays: float64, 3*npts, C-contiguous;
        pinned memory makes the copies overlap the stage kernels.  None drains / discards."""
        def p(a):
            if a is None:
                return None
            assert a.dtype == np.float64 and a.flags.c_contiguous and a.size == 3 * self.npts
            return a.ctypes.data_as(c_dp)
        _chk(self.L.nekcem_b200_step_streamed(self.h, p(hn_in), p(en_in), p(hn_out), p(en_out)))

    def stage(self, rkstep: int):
        _chk(self.L.nekcem_b200_stage(self.h, rkstep))

    def cem_maxwell_op(self, rktime: float):
        """``cem_maxwell_op`` (src/cem_maxwell.F:484-508) as an operator, the way ``amult``
        (:2310-2365) uses it: returns (reshn, resen) after invqmass for the fields on the device;
        the fields themselves are unchanged."""
        _chk(self.L.nekcem_b200_apply_rhs(self.h, float(rktime)))
        return self.get_array("khn"), self.get_array("ken")

    def stage_pack(self, rkstep: int):
        """first half of a stage with option external_exchange: sheet currents + send buffer"""
        _chk(self.L.nekcem_b200_stage_pack(self.h, rkstep))

    def stage_compute(self, rkstep: int):
        """second half: the fused stage on all elements, halo as filled by the caller"""
        _chk(self.L.nekcem_b200_stage_compute(self.h, rkstep))

    def halo_from(self, other: "MaxwellB200"):
        """device copy of ``other``'s send slice for this rank into this context's halo (both
        contexts in this process, same GPU): the tests' stand-in for the NCCL exchange"""
        _chk(self.L.nekcem_b200_halo_exchange_local(self.h, other.h))

    def synchronize(self):
        _chk(self.L.nekcem_b200_synchronize(self.h))

    def geometry_info(self):
        """(elements with constant cofactors, hbm1 == ebm1) found by the setup scan"""
        n, m = C.c_int64(), C.c_int32()
        _chk(self.L.nekcem_b200_geometry_info(self.h, C.byref(n), C.byref(m)))
        return int(n.value), bool(m.value)

    def last_step_ms(self):
        ms = C.c_float(0); n = C.c_int64(0)
        _chk(self.L.nekcem_b200_last_step_ms(self.h, C.byref(ms), C.byref(n)))
        return ms.value, n.value

    def algorithmic_bytes_per_stage(self) -> float:
        b = C.c_double(0)
        _chk(self.L.nekcem_b200_algorithmic_bytes(self.h, C.byref(b)))
        return b.value

    # -- cem_error -----------------------------------------------------------------------
    def error_sums(self, exact_hn, exact_en):
        s = np.zeros(6); m = np.zeros(6)
        _chk(self.L.nekcem_b200_error_sums(
            self.h, _dp(np.ascontiguousarray(exact_hn, dtype=np.float64)),
            _dp(np.ascontiguousarray(exact_en, dtype=np.float64)), _dp(s), _dp(m)))
        return s, m

    def cem_error_mode(self, kind, k, ph, amp, volvm1=None, reduce=None):
        """cem_error against a standing mode evaluated on the device (device-side usersol):
        exact_c = amp[c] * f(kind[c][0], k[0] x + ph[0]) * f(kind[c][1], k[1] y + ph[1])
        * f(kind[c][2], k[2] z + ph[2]), f = 1 / sin / cos for kind 0 / 1 / 2.  Needs the node
        coordinates uploaded as 'xmn','ymn'(,'zmn')."""
        kd = np.ascontiguousarray(np.asarray(kind, dtype=np.int32).reshape(18))
        kk = np.ascontiguousarray(k, dtype=np.float64); pp = np.ascontiguousarray(ph, dtype=np.float64)
        aa = np.ascontiguousarray(amp, dtype=np.float64)
        s, m = np.zeros(6), np.zeros(6)
        _chk(self.L.nekcem_b200_error_sums_mode(self.h, kd.ctypes.data_as(C.POINTER(C.c_int32)),
                                                _dp(kk), _dp(pp), _dp(aa), _dp(s), _dp(m)))
        if reduce is not None:
            s, m = reduce(s, m)
        vol = self.volvm1 if volvm1 is None else volvm1
        l2 = s / vol
        l2 = np.where(l2 > 0, np.sqrt(np.maximum(l2, 0)), l2)
        return l2, m

    def cem_error_planewave(self, omega, k, amp, region, inpml, pml, time, volvm1=None,
                            reduce=None):
        """cem_error against the plane-wave solution of the layered-media tests evaluated on the
        device (device-side usersol): in region r (region[e] in {0,1}),
        exact_c = Re(amp[r][c] * exp(i (k[r] y - omega t) - eta_r pmlfac)); k (2,) and amp (2,6)
        complex; pml = dict(eta, smax, d, y0, sign: (2,) each; order) describes the graded decay
        inside elements with inpml[e] != 0.  Needs 'ymn' uploaded."""
        w = PlaneWave()
        w.omega = float(omega)
        k = np.asarray(k, dtype=np.complex128); amp = np.asarray(amp, dtype=np.complex128)
        for r in range(2):
            w.k_re[r], w.k_im[r] = float(k[r].real), float(k[r].imag)
            for c in range(6):
                w.amp_re[r][c], w.amp_im[r][c] = float(amp[r, c].real), float(amp[r, c].imag)
            for name in ("eta", "smax", "d", "y0", "sign"):
                getattr(w, "pml_" + name)[r] = float(pml[name][r])
        w.pml_order = float(pml["order"])
        reg = np.ascontiguousarray(region, dtype=np.uint8); pm = np.ascontiguousarray(inpml, dtype=np.uint8)
        assert reg.size == self.nelt and pm.size == self.nelt
        s, m = np.zeros(6), np.zeros(6)
        _chk(self.L.nekcem_b200_error_sums_planewave(
            self.h, C.byref(w), reg.ctypes.data_as(C.POINTER(C.c_ubyte)),
            pm.ctypes.data_as(C.POINTER(C.c_ubyte)), float(time), _dp(s), _dp(m)))
        if reduce is not None:
            s, m = reduce(s, m)
        vol = self.volvm1 if volvm1 is None else volvm1
        l2 = s / vol
        l2 = np.where(l2 > 0, np.sqrt(np.maximum(l2, 0)), l2)
        return l2, m

    def cem_error(self, exact_hn, exact_en, volvm1=None, reduce=None):
        """(l2[6], linf[6]) as cem_error; ``reduce(sums, maxes)`` performs the glsc3/glamax
        all-reduce when running on several ranks."""
        s, m = self.error_sums(exact_hn, exact_en)
        if reduce is not None:
            s, m = reduce(s, m)
        vol = self.volvm1 if volvm1 is None else volvm1
        l2 = s / vol
        l2 = np.where(l2 > 0, np.sqrt(np.maximum(l2, 0)), l2)
        return l2, m
