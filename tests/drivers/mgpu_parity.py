"""Multi-GPU parity: run under torchrun with N ranks; every rank takes its pencil-map
partition of a periodic box, steps it with NCCL face exchange, and compares with the oracle
run of the whole mesh.  Prints one line per rank; exits non-zero on mismatch."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
from helpers import arrays_from_refcase, rel_l2, restrict_to_elems  # noqa: E402
from nekcem_b200 import MaxwellB200, comm_unique_id  # noqa: E402
from nekcem_b200.boxcase import gllnid_box  # noqa: E402
from oracle import cases  # noqa: E402

rank, world, local = (int(os.environ[k]) for k in ("RANK", "WORLD_SIZE", "LOCAL_RANK"))
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
case = sys.argv[1] if len(sys.argv) > 1 else "boxper3d"
ade = None
if case in ("boxper3d", "rotated3d"):
    nel, nx1, nsteps = (4, 4, 4 * world), 8, 5
    if case == "rotated3d":
        # every element's local frame turned by one of the 24 proper rotations: the per-peer
        # pack lists must pair face points whose lattices run in different directions
        nx1 = 6
        ref, _ = cases.case_boxper_rotated(nel, nx1, dt=-1e-3)
    else:
        ref = cases.case_boxper(nel, nx1, dt=-1e-3)
    elems = np.nonzero(gllnid_box(*nel, world) == rank)[0]
    s = MaxwellB200(3, nx1, elems.size, device=local, rank=rank, nranks=world)
    s.cem_maxwell_init(arrays_from_refcase(ref, elems))
else:
    # tests/drude or tests/lorentz (2D TE, PML, incident field, ADE) cut into `world` strips of
    # whole element rows along y: exercises the 2D kernel's halo path and every hook across ranks
    ref = getattr(cases, "case_" + case)()
    u = ref.user
    nsteps = 20
    rows = np.arange(ref.nelt) // 4          # .box 4 x 32: element row index
    elems = np.nonzero(rows * world // 32 == rank)[0]
    s = MaxwellB200(2, ref.nx1, elems.size, imode=ref.imode, ifpec=True, ifpml=True, device=local,
                    rank=rank, nranks=world)
    s.cem_maxwell_init(arrays_from_refcase(ref, elems))
    j, amp, phase, omega = u.incident(ref)
    keep, jl = restrict_to_elems(ref, elems, facepts=j)
    if jl.size:
        s.set_incident(jl, amp[:, keep], phase[keep], omega)
    vol = (elems[:, None] * ref.nxyz + np.arange(ref.nxyz)[None, :]).reshape(-1)
    keepn, il = restrict_to_elems(ref, elems, nodes=u.index)
    ncomp = u.jn.size // ref.npts
    npar = u.params.size // ref.npts
    ade = (u.jn.reshape(ncomp, -1)[:, vol].copy(), u.params.reshape(npar, -1)[:, vol].copy(), il)
    if il.size:
        (s.cem_maxwell_drude if case == "drude" else s.cem_maxwell_lorentz)(
            ade[0], None, ade[1], il)
uid = torch.zeros(128, dtype=torch.uint8, device="cuda")
if rank == 0:
    uid.copy_(torch.frombuffer(bytearray(comm_unique_id()), dtype=torch.uint8))
dist.broadcast(uid, 0)
s.comm_init(uid.cpu().numpy().tobytes())
if os.environ.get("NEKCEM_B200_P2P", "1") == "0":
    s.set_option("p2p", 0)   # grouped ncclSend/ncclRecv instead of stores into peer memory
s.setup()
want_transport = "nccl-sendrecv" if os.environ.get("NEKCEM_B200_P2P", "1") == "0" else "peer-memory-push"
assert s.transport() == want_transport, s.transport()
s.set_time(0.0, ref.dt)
s.step(nsteps)
ref.step(nsteps)
vol = (elems[:, None] * ref.nxyz + np.arange(ref.nxyz)[None, :]).reshape(-1)
want = np.concatenate([ref.hn.reshape(3, -1)[:, vol].ravel(), ref.en.reshape(3, -1)[:, vol].ravel()])
got = np.concatenate([s.hn, s.en])
err = rel_l2(got, want)
if ade is not None and ade[2].size:
    jg, _ = s.get_ade()
    ncomp = ref.user.jn.size // ref.npts
    err = max(err, rel_l2(jg, ref.user.jn.reshape(ncomp, -1)[:, vol].ravel()))
vm, peers, nhalo, ni, nb = s.plan()
print(f"rank {rank}/{world}: {s.transport()} rel-L2 vs oracle {err:.3e}; peers {[p for p, _ in peers]} nhalo {nhalo} "
      f"interior {ni} boundary {nb}", flush=True)
ok = torch.tensor([1 if err <= 1e-12 else 0], device="cuda")
dist.all_reduce(ok, op=dist.ReduceOp.MIN)
s.close()
dist.destroy_process_group()
sys.exit(0 if int(ok.item()) == 1 else 1)
