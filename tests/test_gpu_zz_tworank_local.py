"""Two ranks on ONE GPU: the inter-rank machinery (partition, per-peer pack lists, halo slots,
boundary/interior element lists, incident field and graphene sheet current folded into the sent
traces) driven through the transport-independent stage API -- nekcem_b200_stage_pack, a device
copy standing in for the NCCL exchange (nekcem_b200_halo_exchange_local), nekcem_b200_stage_compute
-- and compared with the single-domain oracle.  Covers on a 1-GPU box what tests/test_gpu_multi.py
covers with NCCL on two."""
import numpy as np
import pytest

from helpers import arrays_from_refcase, rel_l2, restrict_to_elems
from nekcem_b200 import MaxwellB200

pytestmark = pytest.mark.gpu
TOL = 1e-12
KTOL = 1e-10  # RK registers of the ill-conditioned sheet ODEs, see tests/test_gpu_zgraphene.py


def _two_ranks(ref, parts, register):
    """contexts for the element partitions ``parts`` (rank r owns parts[r]) of RefCase ``ref``;
    register(s, elems) adds the per-rank hooks before setup"""
    world = len(parts)
    ctx = []
    for r, elems in enumerate(parts):
        s = MaxwellB200(ref.ldim, ref.nx1, elems.size, imode=ref.imode, upwind=True,
                        ifpec=ref.ifpec, ifpml=ref.ifpml, device=0, rank=r, nranks=world)
        s.cem_maxwell_init(arrays_from_refcase(ref, elems))
        register(s, elems)
        s.set_option("external_exchange", 1)
        ctx.append(s)
    ids = [s.face_singletons() for s in ctx]
    counts = np.array([i.size for i in ids])
    all_ids = np.concatenate(ids)
    for s in ctx:
        s.face_remote(counts, all_ids)
        s.setup()
        s.set_time(ref.s.time, ref.s.dt)
    return ctx


def _advance(ctx, ref, nsteps):
    t = ref.s.time
    for _ in range(nsteps):
        for rk in range(1, 6):
            for s in ctx:
                s.stage_pack(rk)
            for d in ctx:
                for o in ctx:
                    if d is not o:
                        d.halo_from(o)
            for s in ctx:
                s.stage_compute(rk)
        t += ref.s.dt
        for s in ctx:
            s.set_time(t, ref.s.dt)
    ref.step(nsteps)


def _field_error(ctx, ref, parts):
    worst = 0.0
    for s, elems in zip(ctx, parts):
        vol = (elems[:, None] * ref.nxyz + np.arange(ref.nxyz)[None, :]).reshape(-1)
        want = np.concatenate([ref.hn.reshape(3, -1)[:, vol].ravel(),
                               ref.en.reshape(3, -1)[:, vol].ravel()])
        worst = max(worst, rel_l2(np.concatenate([s.hn, s.en]), want))
    return worst


def test_two_rank_local_periodic_box():
    """3D periodic box cut into two z-slabs by the reference's pencil map"""
    from nekcem_b200.boxcase import gllnid_box
    from oracle import cases
    nel = (3, 3, 6)
    ref = cases.case_boxper(nel, 6, dt=-1e-3)
    g = gllnid_box(*nel, 2)
    parts = [np.nonzero(g == r)[0] for r in range(2)]
    ctx = _two_ranks(ref, parts, lambda s, e: None)
    vm, peers, nhalo, ni, nb = ctx[0].plan()
    assert nhalo > 0 and nb > 0 and len(peers) == 1
    _advance(ctx, ref, 3)
    assert _field_error(ctx, ref, parts) <= TOL
    for s in ctx:
        s.close()


@pytest.mark.parametrize("which", ["2dgraphene-te", "2dgraphene-tm", "3dgraphene"])
def test_two_rank_local_graphene_sheet_on_the_partition_boundary(which):
    """tests/2dgraphene and tests/3dgraphene cut exactly along the sheet (what the reference's
    own partition of these meshes does at np = 2): each rank owns one side of the sheet, its sheet
    current reaches the other side folded into the packed H trace (H' = H - n x f); the incident
    field of the upper side likewise travels with the trace.  Fields, sheet currents and their RK
    registers against the single-domain oracle."""
    from oracle import cases
    if which == "3dgraphene":
        ref = cases.case_3dgraphene(nx1=6, nel=(3, 6, 3))
        row = (np.arange(ref.nelt) // 3) % 6          # element index along y
        half = 3
    else:
        ref = cases.case_2dgraphene(1 if which.endswith("te") else 2)
        row = np.arange(ref.nelt) // 4
        half = 16
    u = ref.user
    parts = [np.nonzero(row < half)[0], np.nonzero(row >= half)[0]]
    nf, nfp = ref.nxzfl, ref.nxzf * ref.nfaces

    def register(s, elems):
        j, amp, phase, omega = u.incident(ref)
        keep, jl = restrict_to_elems(ref, elems, facepts=j)
        if jl.size:
            s.set_incident(jl, amp[:, keep], phase[keep], omega)
        fac = (elems[:, None] * nfp + np.arange(nfp)[None, :]).reshape(-1)
        keepg, gl = restrict_to_elems(ref, elems, facepts=u.graphindex)
        take = lambda a, m: np.ascontiguousarray(a.reshape(m, nf)[:, fac]).reshape(-1)
        s.cem_graphene_current(take(u.fjn, 18), take(u.kfjn, 18), take(u.graphparams, 12),
                               np.ascontiguousarray(ref.yconduc[fac]), gl)

    ctx = _two_ranks(ref, parts, register)
    assert all(s.plan()[2] > 0 for s in ctx)
    _advance(ctx, ref, 8)
    assert _field_error(ctx, ref, parts) <= TOL
    for s, elems in zip(ctx, parts):
        fac = (elems[:, None] * nfp + np.arange(nfp)[None, :]).reshape(-1)
        fj, kj = s.get_graphene()
        want_f = u.fjn.reshape(18, nf)[:, fac].ravel()
        want_k = u.kfjn.reshape(18, nf)[:, fac].ravel()
        assert np.abs(want_f).max() > 1e-3
        assert rel_l2(fj, want_f) <= TOL and rel_l2(kj, want_k) <= KTOL
    for s in ctx:
        s.close()


def test_two_rank_local_drude():
    """tests/drude cut into two strips: 2D kernel halo path, PML, incident field, ADE"""
    from oracle import cases
    ref = cases.case_drude()
    u = ref.user
    row = np.arange(ref.nelt) // 4
    parts = [np.nonzero(row < 16)[0], np.nonzero(row >= 16)[0]]

    def register(s, elems):
        j, amp, phase, omega = u.incident(ref)
        keep, jl = restrict_to_elems(ref, elems, facepts=j)
        if jl.size:
            s.set_incident(jl, amp[:, keep], phase[keep], omega)
        vol = (elems[:, None] * ref.nxyz + np.arange(ref.nxyz)[None, :]).reshape(-1)
        keepn, il = restrict_to_elems(ref, elems, nodes=u.index)
        if il.size:
            s.cem_maxwell_drude(u.jn.reshape(3, -1)[:, vol].copy(), None,
                                u.params.reshape(2, -1)[:, vol].copy(), il)

    ctx = _two_ranks(ref, parts, register)
    _advance(ctx, ref, 10)
    assert _field_error(ctx, ref, parts) <= TOL
    for s in ctx:
        s.close()
