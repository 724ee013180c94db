"""Multi-GPU parity: run under torchrun with N ranks; every rank takes its pencil-map
partition of a periodic box, steps it with NCCL face exchange, and compares with the oracle
run of the whole mesh.  Prints one line per rank; exits non-zero on mismatch."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
from helpers import arrays_from_refcase, rel_l2  # noqa: E402
from nekcem_b200 import MaxwellB200, comm_unique_id  # noqa: E402
from nekcem_b200.boxcase import gllnid_box  # noqa: E402
from oracle import cases  # noqa: E402

rank, world, local = (int(os.environ[k]) for k in ("RANK", "WORLD_SIZE", "LOCAL_RANK"))
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
nel, nx1, nsteps = (4, 4, 4 * world), 8, 5
ref = cases.case_boxper(nel, nx1, dt=-1e-3)
elems = np.nonzero(gllnid_box(*nel, world) == rank)[0]
s = MaxwellB200(3, nx1, elems.size, device=local, rank=rank, nranks=world)
s.cem_maxwell_init(arrays_from_refcase(ref, elems))
uid = torch.zeros(128, dtype=torch.uint8, device="cuda")
if rank == 0:
    uid.copy_(torch.frombuffer(bytearray(comm_unique_id()), dtype=torch.uint8))
dist.broadcast(uid, 0)
s.comm_init(uid.cpu().numpy().tobytes())
s.setup()
s.set_time(0.0, ref.dt)
s.step(nsteps)
ref.step(nsteps)
vol = (elems[:, None] * ref.nxyz + np.arange(ref.nxyz)[None, :]).reshape(-1)
want = np.concatenate([ref.hn.reshape(3, -1)[:, vol].ravel(), ref.en.reshape(3, -1)[:, vol].ravel()])
got = np.concatenate([s.hn, s.en])
err = rel_l2(got, want)
vm, peers, nhalo, ni, nb = s.plan()
print(f"rank {rank}/{world}: rel-L2 vs oracle {err:.3e}; peers {[p for p, _ in peers]} nhalo {nhalo} "
      f"interior {ni} boundary {nb}", flush=True)
ok = torch.tensor([1 if err <= 1e-12 else 0], device="cuda")
dist.all_reduce(ok, op=dist.ReduceOp.MIN)
s.close()
dist.destroy_process_group()
sys.exit(0 if int(ok.item()) == 1 else 1)
