"""N>1 host logic on CPU: world_size-2 gloo run of the inter-rank face-exchange planning
(partition by the reference's pencil map, singleton-id exchange, per-peer pack lists, halo
slots), checked against the single-domain oracle pairing."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "tests")):
    if p not in sys.path:
        sys.path.insert(0, p)


def _worker(rank, world, port, nel, nx1, q, rotated=False):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from helpers import arrays_from_refcase
    from nekcem_b200 import MaxwellB200
    from nekcem_b200.boxcase import gllnid_box
    from oracle import cases

    # the whole mesh (same on every rank); rotated: every element's local frame turned by one
    # of the 24 proper rotations, so paired face lattices run in different directions
    ref = cases.case_boxper_rotated(nel, nx1)[0] if rotated else cases.case_boxper(nel, nx1)
    rng = np.random.default_rng(7)
    ref.hn[:] = rng.standard_normal(ref.hn.size)  # arbitrary traces
    ref.en[:] = rng.standard_normal(ref.en.size)
    gllnid = gllnid_box(*nel, world)
    elems = np.nonzero(gllnid == rank)[0]
    A = arrays_from_refcase(ref, elems)
    s = MaxwellB200(3, nx1, elems.size, device=-1, rank=rank, nranks=world)
    s.set_faces(A["glo_num"], A["cempec"])
    # singleton exchange over gloo (the library does this over NCCL on GPUs)
    ids = s.face_singletons()
    cnt = torch.tensor([ids.size], dtype=torch.int64)
    cnts = [torch.zeros(1, dtype=torch.int64) for _ in range(world)]
    dist.all_gather(cnts, cnt)
    counts = np.array([int(c.item()) for c in cnts])
    mx = int(counts.max())
    pad = torch.zeros(mx, dtype=torch.int64)
    pad[:ids.size] = torch.from_numpy(ids)
    allp = [torch.zeros(mx, dtype=torch.int64) for _ in range(world)]
    dist.all_gather(allp, pad)
    all_ids = np.concatenate([allp[r].numpy()[:counts[r]] for r in range(world)])
    s.face_remote(counts, all_ids)
    vm, peers, nhalo, ni, nb = s.plan()

    # emulate one exchange: pack own traces at the send face points, swap with the peer
    nxyz, nfp = ref.nxyz, ref.nxzf * ref.nfaces
    loc_cemface = (ref.cemface.reshape(ref.nelt, nfp)[elems] - (elems * nxyz)[:, None]
                   + (np.arange(elems.size) * nxyz)[:, None]).reshape(-1)
    u = np.stack([A["hn"].reshape(3, -1), A["en"].reshape(3, -1)]).reshape(6, -1)
    halo = np.zeros((nhalo, 6))
    off = 0
    for peer, fps in peers:
        send = torch.from_numpy(np.ascontiguousarray(u[:, loc_cemface[fps]].T))
        recv = torch.zeros_like(send)
        reqs = [dist.isend(send, peer), dist.irecv(recv, peer)]
        for r in reqs:
            r.wait()
        halo[off:off + fps.size] = recv.numpy()
        off += fps.size

    # check against the global pairing: for each local face point, the neighbour trace the
    # kernel would read (local node or halo slot) equals the global partner's field values
    order = np.argsort(ref.glo_num, kind="stable")
    partner = np.zeros(ref.nxzfl, dtype=np.int64)
    partner[order[0::2]] = order[1::2]
    partner[order[1::2]] = order[0::2]
    ug = np.stack([ref.hn.reshape(3, -1), ref.en.reshape(3, -1)]).reshape(6, -1)
    glob_fp = (elems[:, None] * nfp + np.arange(nfp)[None, :]).reshape(-1)
    want = ug[:, ref.cemface[partner[glob_fp]]].T
    got = np.where((vm >= 0)[:, None], u[:, np.maximum(vm, 0)].T,
                   halo[np.maximum(-(vm.astype(np.int64) + 3), 0)])
    ok = bool(np.array_equal(want, got)) and vm.max() < elems.size * nxyz and not np.any(vm == -2)
    q.put((rank, ok, len(peers), int(nhalo), ni, nb, int(np.sum(vm <= -3))))
    s.close()
    dist.destroy_process_group()


import pytest  # noqa: E402


@pytest.mark.parametrize("rotated", [False, True])
def test_two_rank_exchange_plan_gloo(rotated):
    nel, nx1, world = (3, 3, 6), 4, 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + (os.getpid() % 2000) + (7 if rotated else 0)
    procs = [ctx.Process(target=_worker, args=(r, world, port, nel, nx1, q, rotated))
             for r in range(world)]
    for p in procs:
        p.start()
    res = [q.get(timeout=300) for _ in range(world)]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    for rank, ok, npeers, nhalo, ni, nb, nremote in sorted(res):
        assert ok, f"rank {rank}: neighbour traces differ from the single-domain pairing"
        assert npeers == 1
        # z-slabs of 3 layers, periodic in z: both slab faces are remote: 2*3*3 faces * 16 pts
        assert nhalo == 2 * 9 * 16 and nremote == nhalo
        assert ni == 9 and nb == 18


def _worker_2d_sheet(rank, world, port, q):
    """tests/2dgraphene cut along the sheet (rows 0-15 / 16-31), as the reference's partition of
    this mesh does at np = 2: 2D face numbering across ranks, and the registration of a sheet
    whose every point lies on an inter-rank face (host side of nekcem_b200_set_graphene)."""
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from helpers import arrays_from_refcase, restrict_to_elems
    from nekcem_b200 import MaxwellB200
    from oracle import cases

    ref = cases.case_2dgraphene(1)
    u = ref.user
    row = np.arange(ref.nelt) // 4
    elems = np.nonzero((row >= 16) == (rank == 1))[0]
    A = arrays_from_refcase(ref, elems)
    s = MaxwellB200(2, ref.nx1, elems.size, imode=1, ifpml=True, device=-1, rank=rank, nranks=world)
    s.set_faces(A["glo_num"], A["cempec"])
    nf, nfp = ref.nxzfl, ref.nxzf * ref.nfaces
    fac = (elems[:, None] * nfp + np.arange(nfp)[None, :]).reshape(-1)
    keep, gl = restrict_to_elems(ref, elems, facepts=u.graphindex)
    take = lambda a, m: np.ascontiguousarray(a.reshape(m, nf)[:, fac]).reshape(-1)
    s.cem_graphene_current(take(u.fjn, 18), take(u.kfjn, 18), take(u.graphparams, 12),
                           np.ascontiguousarray(ref.yconduc[fac]), gl)
    ids = s.face_singletons()
    cnt = torch.tensor([ids.size], dtype=torch.int64)
    cnts = [torch.zeros(1, dtype=torch.int64) for _ in range(world)]
    dist.all_gather(cnts, cnt)
    counts = np.array([int(c.item()) for c in cnts])
    mx = int(counts.max())
    pad = torch.zeros(mx, dtype=torch.int64)
    pad[:ids.size] = torch.from_numpy(ids)
    allp = [torch.zeros(mx, dtype=torch.int64) for _ in range(world)]
    dist.all_gather(allp, pad)
    all_ids = np.concatenate([allp[r].numpy()[:counts[r]] for r in range(world)])
    s.face_remote(counts, all_ids)
    vm, peers, nhalo, ni, nb = s.plan()
    # the peer's send list pairs point by point with ours: exchange the y coordinate of the send
    # points (must all be 0: the sheet) and the x coordinate (must agree pairwise)
    cf = ref.cemface[fac]                                   # global volume node per local face pt
    fps = peers[0][1]
    send = torch.from_numpy(np.ascontiguousarray(np.stack([ref.xm1[cf[fps]], ref.ym1[cf[fps]]], 1)))
    recv = torch.zeros_like(send)
    reqs = [dist.isend(send, peers[0][0]), dist.irecv(recv, peers[0][0])]
    for r in reqs:
        r.wait()
    ok = (np.abs(send[:, 1].numpy()).max() < 1e-12 and np.abs(recv[:, 1].numpy()).max() < 1e-12
          and np.abs(send[:, 0].numpy() - recv[:, 0].numpy()).max() < 1e-12)
    on_halo = int(np.sum(vm[gl] <= -3))
    q.put((rank, bool(ok), len(peers), int(nhalo), ni, nb, on_halo, int(gl.size)))
    s.close()
    dist.destroy_process_group()


def test_two_rank_2d_sheet_on_the_partition_boundary_gloo():
    world = 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + (os.getpid() % 2000) + 23
    procs = [ctx.Process(target=_worker_2d_sheet, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = [q.get(timeout=300) for _ in range(world)]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    for rank, ok, npeers, nhalo, ni, nb, on_halo, ng in sorted(res):
        assert ok, f"rank {rank}: send lists of the two ranks do not pair point by point"
        assert npeers == 1 and nhalo == 4 * 9        # one row of 4 element faces, 9 points each
        assert ni == 60 and nb == 4
        assert ng == 36 and on_halo == 36            # the whole sheet of this rank is remote
