"""The per-face-point graphene update of the CUDA library (nekcem_b200/csrc/stage_graphene.h, a
__host__ __device__ function) compiled with g++ and compared with the oracle's restatement of
cem_3d/te/tm_graphene_current (src/cem_maxwell.F:2827-3093) on the face values of a running
graphene case -- CPU only: checks the arithmetic the GPU kernel executes, not the kernel."""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

from oracle import cases

HERE = os.path.dirname(os.path.abspath(__file__))


@pytest.fixture(scope="module")
def harness(tmp_path_factory):
    out = str(tmp_path_factory.mktemp("gp") / "libgp.so")
    subprocess.check_call(["g++", "-O2", "-ffp-contract=off", "-shared", "-fPIC", "-o", out,
                           os.path.join(HERE, "graphene_point_harness.cpp")])
    L = C.CDLL(out)
    dp = C.POINTER(C.c_double)
    L.graphene_points.argtypes = [C.c_int, C.c_int] + [dp] * 8 + [C.c_double] * 3
    return L


@pytest.mark.parametrize("which", ["3d", "te", "tm"])
def test_graphene_point_matches_oracle(harness, which):
    c = (cases.case_3dgraphene(nel=(3, 12, 3)) if which == "3d"
         else cases.case_2dgraphene(1 if which == "te" else 2))
    u = c.user
    nf, j = c.nxzfl, u.graphindex
    ng = j.size
    dp = lambda a: a.ctypes.data_as(C.POINTER(C.c_double))
    seen = []

    orig = u.userfsrc(c)

    def spy(tt, *src):
        # face values exactly as the reference's userfsrc sees them (after userinc)
        H = np.ascontiguousarray(c.fhn.reshape(3, nf)[:, j])
        E = np.ascontiguousarray(c.fen.reshape(3, nf)[:, j])
        nrm = np.ascontiguousarray(np.stack([c.unxm[j], c.unym[j], c.unzm[j]]))
        Yfac = 0.5 / c.Y_0[j]
        yc = np.ascontiguousarray(c.yconduc[j])
        par = np.ascontiguousarray(u.graphparams.reshape(12, nf)[:, j])
        fj = np.ascontiguousarray(u.fjn.reshape(18, nf)[:, j])
        kj = np.ascontiguousarray(u.kfjn.reshape(18, nf)[:, j])
        k = c.s.rkstep - 1
        harness.graphene_points(c.imode, ng, dp(H), dp(E), dp(nrm), dp(Yfac), dp(yc), dp(par),
                                dp(fj), dp(kj), c.s.rk4a[k], c.s.rk4b[k], c.s.dt)
        orig(tt, *src)
        fo = u.fjn.reshape(18, nf)[:, j]
        ko = u.kfjn.reshape(18, nf)[:, j]
        seen.append((np.abs(fj - fo).max(), np.abs(ko - kj).max(), np.abs(fo).max()))
        assert np.array_equal(fj, fo) and np.array_equal(kj, ko)

    c.set_callback("userfsrc", spy)
    c.step(4)
    assert len(seen) == 20 and max(s[2] for s in seen) > 1e-3
