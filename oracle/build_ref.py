"""Recipe for oracle/_ref/libnekcem_ref.so -- the REFERENCE'S OWN hot path, executable here.

TEST INFRASTRUCTURE ONLY.  The reference's Maxwell path is Fortran + one C library; this image
has gcc but no Fortran compiler.  The recipe therefore

  1. translates the reference's own Fortran routines on the path, read from
     /root/reference/src where they lie, into C with oracle/f2c_lite.py (mechanical,
     statement by statement, same arithmetic order; see that file's header),
  2. compiles the translation together with the reference's own gather-scatter library
     (src/jl/gs.c and its dependencies, compiled unchanged, single process, no MPI) and the
     small harness oracle/ref_harness.c,
  3. writes everything to oracle/_ref/ (git-ignored; the .so travels to the GPU box).

Preprocessor configuration: -DMPI -DMPIIO -DMAXWELL as bin/configurenek sets them, plus
-DNOTIMER (a switch of the reference itself, src/nek5_mxm_wrapper.F:16: it removes only the
flop counters of mxm).  `real` is 8 bytes as with -fdefault-real-8 (bin/configurenek:117).

Run:  python oracle/build_ref.py          (needs /root/reference; a no-op without it)
"""
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
REF = os.environ.get("NEKCEM_REFERENCE", "/root/reference")
OUT = os.path.join(HERE, "_ref")
LIB = os.path.join(OUT, "libnekcem_ref.so")

# reference file -> the program units on the path
UNITS = [
    ("src/cem_maxwell.F", [
        "cem_maxwell_op_rk", "cem_maxwell_op", "cem_maxwell", "maxwell_wght_curl",
        "cem_maxwell_restrict_to_face", "cem_maxwell_flux", "cem_maxwell_flux2d",
        "cem_maxwell_flux3d", "cem_maxwell_flux_pec", "cem_maxwell_add_flux_to_res",
        "cem_maxwell_invqmass", "rk_maxwell_ab", "cem_maxwell_drude", "cem_maxwell_lorentz",
        "cem_maxwell_materials", "cem_maxwell_pec_init",
        # graphene sheets: surface-current ADEs called from the .usr's userfsrc (8f rank 1/4)
        "cem_3d_graphene_current", "cem_te_graphene_current", "cem_tm_graphene_current"]),
    ("src/cem_maxwell_pml.F", ["pml_step", "pml_faces", "march_faces", "pml_fill_faceary",
                               "dir_local_to_global", "pml_extent_and_tags", "pml_calc_sigma"]),
    ("src/cem_common.F", ["rk_c", "rk4_upd", "rk_storage", "cem_set_fc_ptr", "cem_error"]),
    ("src/nek5_courant.F", ["get_dxmin"]),
    ("src/nek5_grad.F", ["local_grad3", "local_grad2"]),
    ("src/nek5_mxm_wrapper.F", ["mxm"]),
    ("src/nek5_mxm_std.F", ["mxmf2"] + ["mxf%d" % k for k in range(1, 25)] + ["mxm44_0"]),
    ("src/nek5_mat1.F", ["chsign", "rzero", "rone", "copy", "addcol3", "addcol4", "subcol3",
                         "subcol4", "ascol5", "col2", "col3", "invcol1", "invcol3", "invers2",
                         "vdot2", "vdot3", "vcross", "unitvec", "rzero3", "cmult", "sub3",
                         "izero", "vlmax", "vlmin", "glsc3", "glamax", "glmin", "glmax",
                         "addtnsr", "mod1"]),
    # setup routines whose OUTPUT the path consumes: GLL nodes/weights and
    # derivative matrix, metric cofactors / Jacobian / mass, face areas and normals
    ("src/nek5_speclib.F", ["zwgll", "zwglj", "zwgljd", "jacg", "jacobf", "zwgjd", "endw1",
                            "endw2", "gammaf", "pnormj", "dgll", "pnleg", "pndleg"]),
    ("src/nek5_coef.F", ["xyzrst", "glmapm1", "chkjac", "geodat1", "setarea", "area2", "area3",
                         "setwgtr", "set_unr"]),
    ("src/nek5_subs2.F", ["facexv", "setaxdy", "setaxw1"]),
    # node coordinates from the element vertices, incl. circular-arc sides (tests/cylwave); the
    # other curved-side generators sphsrf / gensrf are outside the configs of the path
    ("src/nek5_genxyz.F", ["genxyz", "setzgml", "arcsrf"]),
    # setup routines that define the face numbering the path relies on
    ("src/nek5_connect11.F", ["initds", "dsset"]),
    # the analytic solutions the reference's own tests check against: the
    # usersol of each shipped case, emitted as usersol__<case>_
    ("tests/3dboxper/3dboxper.usr", ["usersol"], "__3dboxper"),
    ("tests/3dboxpec/3dboxpec.usr", ["usersol"], "__3dboxpec"),
    ("tests/2dboxper/2dboxper.usr", ["usersol"], "__2dboxper"),
    ("tests/2dboxpec/2dboxpec.usr", ["usersol"], "__2dboxpec"),
    # the other callbacks of the shipped cases (drude.usr / lorentz.usr evaluate their analytic
    # solution in COMPLEX arithmetic: C99 double _Complex, compiled with -fcx-fortran-rules)
    ("tests/3dboxper/3dboxper.usr", ["usrdat2"], "__3dboxper"),
    ("tests/3dboxpec/3dboxpec.usr", ["usrdat2"], "__3dboxpec"),
    ("tests/3ddielectric/3ddielectric.usr", ["userinc", "usersol", "userini", "uservp",
                                             "usrdat2"], "__3ddielectric"),
    ("tests/3dboxpml/3dboxpml.usr", ["usersrc", "usrdat2"], "__3dboxpml"),
    ("tests/drude/drude.usr", ["userinc", "usersrc", "usersol", "userini", "uservp"], "__drude"),
    ("tests/lorentz/lorentz.usr", ["userinc", "usersrc", "usersol", "userini", "uservp"],
     "__lorentz"),
    # the optional modal filter at the end of a time step (src/cem_maxwell.F:342, param(18) = 1):
    # its two workers.  q_filter itself (six filterq calls around a SAVEd matrix, which the
    # translator cannot keep across calls for run-time array sizes) is driven by the tests
    ("src/nek5_filter.F", ["filterq", "build_new_filter"]),
    ("src/nek5_grad.F", ["legendre_poly"]),
    ("src/nek5_mat1.F", ["ident", "transpose", "gaujordf", "vlamax"]),
    # output hand-off: the per-node interleave of cem_out (src/io.F:207-210)
    ("src/io_dumpvtk.F", ["vtk_nonswap_field"]),
    ("tests/cylwave/cylwave.usr", ["usrdat", "usrdat2"], "__cylwave"),
    ("src/cem_common.F", ["geom_xyradius"]),
    ("tests/2ddielectric/2ddielectric.usr", ["userinc", "usersol", "userini", "uservp",
                                             "usrdat2"], "__2ddielectric"),
    ("tests/2dboxpml/2dboxpml.usr", ["usersrc", "usrdat2"], "__2dboxpml"),
    ("tests/3dgraphene/3dgraphene.usr", ["userinc", "userfsrc", "usersol", "userini", "uservp",
                                         "usrdat2"], "__3dgraphene"),
    ("tests/2dgraphene/2dgraphene.usr", ["userinc", "userfsrc", "usersol", "userini", "uservp",
                                         "usrdat2"], "__2dgraphene"),
]
# reference gather-scatter library, compiled unchanged (flags of bin/configurenek:132-139
# without -DMPI: single process)
JL = ["gs.c", "gs_local.c", "comm.c", "crystal.c", "sarray_transfer.c", "sarray_sort.c",
      "sort.c", "fail.c", "tensor.c"]
JL_FLAGS = ["-DUNDERSCORE", "-DGLOBAL_LONG_LONG", "-DUSE_NAIVE_BLAS"]
DEFINES = ("MPI", "MPIIO", "MAXWELL", "NOTIMER")

# Drop-in variant: the same translated reference + THIS REPO'S fixed-form shim
# (fortran/cem_maxwell_b200_f77.F, translated by the same tool) linked against the product
# library.  cem_maxwell_drude / cem_maxwell_lorentz and the three graphene-current routines are
# left out so that the .usr's calls resolve to the library's twins, exactly as a -DB200 build of
# the reference would link.
LIB_DROPIN = os.path.join(OUT, "libnekcem_ref_dropin.so")
REPO = os.path.dirname(HERE)
SHIM = (os.path.join(REPO, "fortran", "cem_maxwell_b200_f77.F"),
        ["b200_copy_all_in", "b200_update_device", "b200_op_rk", "b200_update_host",
         "b200_restart_out", "b200_restart_swap", "b200_copy_all_out"])
PRODUCT_LIB_DIR = os.path.join(REPO, "nekcem_b200", "lib")


def available() -> bool:
    return os.path.isdir(os.path.join(REF, "src"))


def _f2c_lite():
    """the translator module, loaded by file path: oracle/ itself must never be put on sys.path
    (`import oracle` would then resolve to oracle/oracle.py instead of the package, also in
    processes spawned later)"""
    import importlib.util
    name = "oracle_f2c_lite"
    if name in sys.modules:
        return sys.modules[name]
    spec = importlib.util.spec_from_file_location(name, os.path.join(HERE, "f2c_lite.py"))
    mod = importlib.util.module_from_spec(spec)
    sys.modules[name] = mod
    spec.loader.exec_module(mod)
    return mod


def build(force: bool = False, verbose: bool = False) -> str | None:
    """returns the library path, or None when the reference tree is absent and no prebuilt
    library exists"""
    if not available():
        return LIB if os.path.exists(LIB) else None
    srcs = [os.path.join(REF, u[0]) for u in UNITS] + [os.path.join(REF, "src/jl", f) for f in JL]
    mine = [os.path.join(HERE, f) for f in ("f2c_lite.py", "build_ref.py", "ref_harness.c")]
    mine += [SHIM[0], os.path.join(REPO, "fortran", "NEKCEM_B200")]
    if not force and os.path.exists(LIB):
        t = os.path.getmtime(LIB)
        if all(os.path.getmtime(s) < t for s in srcs + mine):
            return LIB
    f2c_lite = _f2c_lite()
    os.makedirs(OUT, exist_ok=True)
    inc = [os.path.join(REF, "tests/3dboxper"), os.path.join(REF, "src")]
    ctext, em = f2c_lite.translate([(os.path.join(REF, u[0]),) + tuple(u[1:]) for u in UNITS], inc,
                                   DEFINES)
    gen = os.path.join(OUT, "ref_gen.c")
    with open(gen, "w") as f:
        f.write(ctext)
    cmd = (["gcc", "-O3", "-std=gnu11", "-ffp-contract=off", "-fcx-fortran-rules", "-fPIC",
            "-shared", "-w", "-I" + os.path.join(REF, "src/jl")] + JL_FLAGS +
           ["-o", LIB, gen, os.path.join(HERE, "ref_harness.c")] +
           [os.path.join(REF, "src/jl", f) for f in JL] +
           # byte-order helpers of the reference's VTK writer (swap_float_byte, ...), unchanged
           [os.path.join(REF, "src", "io_util.c"), "-I" + os.path.join(REF, "src")] + ["-lm"])
    if verbose:
        print(" ".join(cmd))
    subprocess.check_call(cmd)
    build_dropin(inc, verbose)
    return LIB


def build_dropin(inc, verbose=False):
    """libnekcem_ref_dropin.so (needs nekcem_b200/lib/libnekcem_b200.so; skipped without it)"""
    f2c_lite = _f2c_lite()
    if not os.path.exists(os.path.join(PRODUCT_LIB_DIR, "libnekcem_b200.so")):
        return None
    units = []
    for u in UNITS:
        names = [n for n in u[1] if n not in (
            "cem_maxwell_drude", "cem_maxwell_lorentz", "cem_3d_graphene_current",
            "cem_te_graphene_current", "cem_tm_graphene_current")]
        units.append((os.path.join(REF, u[0]), names) + tuple(u[2:]))
    units.append(SHIM)
    ctext, em = f2c_lite.translate(units, inc + [os.path.join(REPO, "fortran")], DEFINES)
    gen = os.path.join(OUT, "ref_dropin_gen.c")
    with open(gen, "w") as f:
        f.write(ctext)
    cmd = (["gcc", "-O3", "-std=gnu11", "-ffp-contract=off", "-fcx-fortran-rules", "-fPIC",
            "-shared", "-w", "-I" + os.path.join(REF, "src/jl")] + JL_FLAGS +
           ["-o", LIB_DROPIN, gen, os.path.join(HERE, "ref_harness.c")] +
           [os.path.join(REF, "src/jl", f) for f in JL] +
           ["-L" + PRODUCT_LIB_DIR, "-lnekcem_b200",
            "-Wl,-rpath,$ORIGIN/../../nekcem_b200/lib", "-lm"])
    if verbose:
        print(" ".join(cmd))
    subprocess.check_call(cmd)
    return LIB_DROPIN


if __name__ == "__main__":
    p = build(force=True, verbose=True)
    print("built" if p else "reference tree absent", p)
