"""GPU parity of the plane-sweep stage kernel (nekcem_b200/csrc/stage_sweep.cu, nx1 = 11..16,
option `sweep`) through the C ABI: against the oracle (<= 1e-12 rel-L2) and against the slab
kernel (same products, lifts added after the whole volume curl: rounding-level differences).
`pipeline_ctas` caps the persistent grid so that every CTA walks through many work items on these
small meshes (the ring of bulk copies runs across item boundaries and the mbarrier phases wrap)."""
import numpy as np
import pytest

from helpers import incident_3ddielectric, rel_l2, solver_from_refcase

pytestmark = pytest.mark.gpu
TOL = 1e-12
SWEEP_ORDERS = [11, 12, 13, 14, 15, 16]


def _fields(s):
    if hasattr(s, "get_array"):
        return np.concatenate([s.get_array("hn"), s.get_array("en")])
    return np.concatenate([s.hn, s.en])


def _state(s):
    return np.concatenate([s.get_array(k) for k in ("hn", "en", "khn", "ken")])


@pytest.mark.parametrize("ctas", [0, 2, 6])
@pytest.mark.parametrize("nx1", SWEEP_ORDERS)
def test_sweep_agrees_with_slab_kernel(nx1, ctas):
    from nekcem_b200 import MaxwellB200
    from nekcem_b200.boxcase import BoxCase
    case = BoxCase((4, 3, 3), nx1)
    out = []
    for sweep in (0, 1):
        s = MaxwellB200(3, nx1, case.nelt, device=0)
        s.cem_maxwell_init(case.arrays())
        s.set_option("const_metrics", 0)
        s.set_option("sweep", sweep)
        s.set_option("pipeline_ctas", ctas)
        s.setup()
        s.set_time(0.0, 1e-3)
        s.step(2)
        out.append(_state(s))
        s.close()
    assert rel_l2(out[0], out[1]) <= 1e-13


@pytest.mark.parametrize("nx1", SWEEP_ORDERS)
def test_sweep_periodic_box_vs_oracle(nx1):
    from oracle import cases
    c = cases.case_boxper((3, 3, 4), nx1, dt=-1e-3)
    s = solver_from_refcase(c)
    s.set_option("const_metrics", 0)
    s.set_option("sweep", 1)
    s.set_option("pipeline_ctas", 4)
    c.step(2); s.step(2)
    assert rel_l2(_fields(s), _fields(c)) <= TOL
    # the RK register is dt*res: a difference of O(N^2) larger terms (see test_gpu_pipe.py)
    kg = np.concatenate([s.get_array("khn"), s.get_array("ken")])
    assert rel_l2(kg, np.concatenate([c.khn, c.ken])) <= 1e-10
    s.close()


@pytest.mark.parametrize("nx1", [11, 13, 16])
def test_sweep_pec_cavity_deformed_elements(nx1):
    """PEC walls and non-affine elements: per-node cofactors really vary inside an element"""
    from oracle import cases, oracle as O
    mesh = O.box_mesh((2, 2, 3), ((-1.0, 1.0),) * 3, ("PEC",) * 6)

    def warp(case):
        x, y, z = case.xm1.copy(), case.ym1.copy(), case.zm1.copy()
        case.xm1[:] = x + 0.08 * np.sin(np.pi * y) * np.sin(np.pi * z)
        case.ym1[:] = y + 0.06 * np.sin(np.pi * x) * np.sin(np.pi * z)
        case.zm1[:] = z + 0.05 * np.sin(np.pi * x) * np.sin(np.pi * y)

    c = O.RefCase(mesh, nx1 - 1, upwind=True, usrdat2=warp)
    c.set_dt(-2e-3)
    c.hn[:], c.en[:] = cases.usersol_3dboxpec(c, 0.0)
    assert np.abs(c.rymn).max() > 1e-3  # genuinely curved metrics
    s = solver_from_refcase(c)
    s.set_option("sweep", 1)
    s.set_option("pipeline_ctas", 2)
    c.step(2); s.step(2)
    assert rel_l2(_fields(s), _fields(c)) <= TOL
    s.close()
