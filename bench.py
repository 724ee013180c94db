#!/usr/bin/env python
"""bench.py -- GDOF-RK-stage/s (FP64) of the fused Maxwell RK step on B200.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--elems E] [--order P]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...
    python bench.py --impl reference ...      (CPU arm: the reference's own path, oracle/_ref,
                                               on all host cores; the oracle port if absent)

Workload (BASELINE.json configs[4]): synthetic 3D periodic box, 64^3 hex elements at N=7 per
GPU (weak scaling: the global box is 64 x 64 x 64*N elements, split into z-slabs by the
reference's pencil map), tests/3dboxper initial condition, upwind flux, RK45.  A "step" is
one time step = 5 RK stages over every node.  1 DOF = 1 grid node carrying 6 components
(SURVEY.md 8d).  One JSON line is printed by rank 0.
"""
from __future__ import annotations

import argparse
import json
import math
import os
import subprocess
import sys
import tempfile
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

# stdout carries exactly ONE JSON line: native libraries (NCCL prints "NCCL version ..." on
# stdout when NCCL_DEBUG is set) and anything else that writes to fd 1 are sent to stderr
_JSON_OUT = os.fdopen(os.dup(1), "w")
os.dup2(2, 1)
sys.stdout = sys.stderr

METRIC = "GDOF-RK-stage/s (FP64)"
UNIT = "Gnode-stage/s"


def measured_peak():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md)"


# ------------------------------------------------------------------------------------------
# CPU arm: the oracle port (the reference itself cannot be built: no Fortran/MPI toolchain)
# ------------------------------------------------------------------------------------------
def cpu_oracle_rate(elems: int, nx1: int, steps: int, warmup: int):
    """Times the oracle (oracle/nekcem_oracle.c, OpenMP over the host cores) on a bounded
    sample of the same workload: periodic box elems^3, order nx1-1."""
    from oracle import cases, oracle as O
    L = O.lib()
    c = cases.case_boxper((elems,) * 3, nx1)
    threads = int(L.ora_num_threads())
    c.step(max(warmup, 1))
    t0 = time.perf_counter()
    c.step(steps)
    dt = time.perf_counter() - t0
    rate = c.npts * 5.0 * steps / dt / 1e9
    return rate, threads, dt, c.npts


def _ref_worker(elems: int, nx1: int, steps: int, warmup: int, sync_dir: str, idx: int):
    """one replica of the translated reference (oracle/_ref): setup, warm-up, file barrier,
    then `steps` timed calls of the reference's cem_maxwell_op_rk"""
    os.environ["OMP_NUM_THREADS"] = "1"
    from oracle import cases, refrun
    c = cases.case_boxper((elems,) * 3, nx1)
    r = refrun.ReferenceRun(c)
    r.step(max(warmup, 1))
    open(os.path.join(sync_dir, f"ready_{idx}"), "w").close()
    go = os.path.join(sync_dir, "go")
    while not os.path.exists(go):
        time.sleep(0.005)
    t0 = time.perf_counter()
    r.step(steps)
    dt = time.perf_counter() - t0
    print(json.dumps({"dt": dt, "npts": int(c.npts)}), file=_JSON_OUT, flush=True)


def cpu_reference_rate(elems: int, nx1: int, steps: int, warmup: int):
    """The reference's own hot path (oracle/_ref: its Fortran translated to C + its src/jl gs
    library) on all host cores.  This image has no MPI, so the cores are filled the way
    `mpiexec -np P` would fill them but without the inter-rank exchange: P independent
    single-process replicas, each advancing its own periodic box of elems^3 elements
    (an upper bound on what the MPI reference could reach on the same cores).
    Returns None when oracle/_ref is not available."""
    from oracle import refrun
    if not refrun.available():
        return None
    procs = len(os.sched_getaffinity(0))
    sync_dir = tempfile.mkdtemp(prefix="nekcem_ref_")
    ws = [subprocess.Popen([sys.executable, os.path.abspath(__file__), "--ref-worker",
                            f"{elems},{nx1},{steps},{warmup},{sync_dir},{i}"],
                           stdout=subprocess.PIPE, text=True) for i in range(procs)]
    while sum(os.path.exists(os.path.join(sync_dir, f"ready_{i}")) for i in range(procs)) < procs:
        if any(w.poll() not in (None, 0) for w in ws):
            for w in ws:
                w.kill()
            return None
        time.sleep(0.01)
    open(os.path.join(sync_dir, "go"), "w").close()
    outs = [json.loads(w.communicate()[0].strip().splitlines()[-1]) for w in ws]
    import shutil
    shutil.rmtree(sync_dir, ignore_errors=True)
    dt = max(o["dt"] for o in outs)
    npts = sum(o["npts"] for o in outs)
    return npts * 5.0 * steps / dt / 1e9, procs, dt, npts


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    nx1 = args.order + 1
    ref = cpu_reference_rate(args.ref_elems, nx1, args.steps, args.warmup)
    if ref is not None:
        rate, threads, dt, npts = ref
        kind = "reference"
        sample = (f"{threads} independent single-process replicas of the reference's own path "
                  f"(oracle/_ref: its Fortran translated to C + its src/jl gs library; no MPI in "
                  f"this image), each a periodic box of {args.ref_elems}^3 elements at "
                  f"N={args.order} ({npts} nodes in total), {args.steps} steps per run")
    else:
        elems = args.cpu_elems
        rate, threads, dt, npts = cpu_oracle_rate(elems, nx1, args.steps, args.warmup)
        kind = "port"
        sample = (f"periodic box {elems}^3 elements, N={args.order} ({npts} nodes), "
                  f"{args.steps} steps per run; oracle port (oracle/_ref did not travel)")
    line = {
        "impl": "reference", "metric": METRIC, "value": rate, "unit": UNIT, "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * dt / args.steps,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64",
        "data": "synthetic",
        "config": {"workload": f"synthetic 3D periodic box, 64^3 hex elements/GPU at N={args.order}"
                               " (CPU arm runs a bounded sample of it)",
                   "sample": sample},
        "cpu_baseline": {"value": rate, "unit": UNIT, "cores": threads, "kind": kind,
                         "sample": sample},
        "e2e": {"value": rate, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), file=_JSON_OUT, flush=True)


# ------------------------------------------------------------------------------------------
class ClockSampler:
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
         "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index: int):
        self.f = tempfile.NamedTemporaryFile("w+", suffix=".csv", delete=False)
        self.p = None
        try:
            self.p = subprocess.Popen(
                ["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                 "-lms", "200", "-i", str(gpu_index)], stdout=self.f, stderr=subprocess.DEVNULL)
        except Exception:
            self.p = None

    def stop(self):
        out = {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        if self.p is None:
            return out
        time.sleep(0.25)
        self.p.terminate()
        try:
            self.p.wait(timeout=5)
        except Exception:
            self.p.kill()
        self.f.flush()
        rows = [l.strip().split(",") for l in open(self.f.name) if l.strip()]
        os.unlink(self.f.name)
        sm, smax, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in rows:
            if len(r) < 9:
                continue
            try:
                sm.append(float(r[1])); smax.append(float(r[2]))
            except ValueError:
                continue
            for nm, v in zip(names, r[5:9]):
                if v.strip().lower() == "active":
                    reasons.add(nm)
        if sm:
            out.update(sm_mhz=float(np.median(sm)), sm_max_mhz=float(max(smax)),
                       reasons=sorted(reasons), samples=len(sm))
        return out


def run_gpu(args):
    import torch
    import torch.distributed as dist

    from nekcem_b200 import MaxwellB200, comm_unique_id
    from nekcem_b200.boxcase import BoxCase

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world != args.gpus:
        if world == 1 and args.gpus > 1:
            raise SystemExit("launch with torch.distributed.run --nproc-per-node N for --gpus N")
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device (B200); there is no CPU fallback")
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    E, nx1 = args.elems, args.order + 1
    strong = args.scaling == "strong"
    # weak scaling (default): one E^3 slab per GPU; strong: the global E^3 box split over the GPUs
    nel = (E, E, E) if strong else (E, E, E * world)
    case = BoxCase(nel, nx1, rank=rank, nranks=world, length=2 * math.pi)
    slv = MaxwellB200(3, nx1, case.nelt, device=local, rank=rank, nranks=world)
    t_setup = time.perf_counter()
    slv.cem_maxwell_init(case.lazy(), free_after_upload=True)
    if world > 1:
        uid = torch.zeros(128, dtype=torch.uint8, device="cuda")
        if rank == 0:
            uid.copy_(torch.frombuffer(bytearray(comm_unique_id()), dtype=torch.uint8))
        dist.broadcast(uid, 0)
        slv.comm_init(bytes(uid.cpu().numpy().tobytes()))
    slv.setup()
    if args.metrics == "stream":
        slv.set_option("const_metrics", 0)
    # CFL-limited dt as in the synthetic .rea (param(12)=+0.1 -> dt = 0.1*dxmin)
    from nekcem_b200.boxcase import gll
    z, _ = gll(nx1)
    dxmin = 0.5 * min(case.h) * 0.5 * float(np.min(z[2:] - z[:-2])) if nx1 > 2 else min(case.h)
    dt = 0.1 * dxmin
    slv.set_time(0.0, dt)
    t_setup = time.perf_counter() - t_setup
    npts_global = case.npts * world

    # ---- device-resident timing ---------------------------------------------------------
    W, K = max(args.warmup, 3), args.steps
    slv.step(W)
    barrier()
    sampler = ClockSampler(local) if rank == 0 else None
    barrier()
    slv.cem_maxwell_op_rk(K, sync=False)
    slv.synchronize()
    barrier()
    ms, launches = slv.last_step_ms()
    clocks = sampler.stop() if sampler else None
    tms = torch.tensor([ms], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(tms, op=dist.ReduceOp.MAX)
    ms_max = float(tms.item())
    value = npts_global * 5.0 * K / (ms_max * 1e-3) / 1e9

    # analytic-solution check of the timed state (the reference's userchk, 3dboxper.usr:169-216)
    tnow = slv.time
    shn, sen = case.fields(tnow)
    s, m = slv.error_sums(shn, sen)
    del shn, sen
    red = torch.tensor(np.concatenate([s, m]), dtype=torch.float64, device="cuda")
    if world > 1:
        sums = red[:6].clone(); mx = red[6:].clone()
        dist.all_reduce(sums, op=dist.ReduceOp.SUM)
        dist.all_reduce(mx, op=dist.ReduceOp.MAX)
        red = torch.cat([sums, mx])
    red = red.cpu().numpy()
    l2 = np.sqrt(red[:6] / (case.volume_global * 1.0))
    linf = red[6:]

    # ---- end-to-end through the C ABI with HOST buffers ------------------------------------
    # every step: H2D of HN,EN from pinned host memory, one time step, D2H of HN,EN
    # (the `!$ACC UPDATE DEVICE/HOST(hn,en)` seams of the reference, drude.usr:94, 3dboxper.usr:199)
    Ke = max(1, min(args.e2e_steps, K))
    n3 = 3 * case.npts
    pin_h = torch.empty(n3, dtype=torch.float64, pin_memory=True)
    pin_e = torch.empty(n3, dtype=torch.float64, pin_memory=True)
    hn_host, en_host = pin_h.numpy(), pin_e.numpy()
    hn_host[:] = slv.hn
    en_host[:] = slv.en
    import ctypes as C
    from nekcem_b200.api import ARRAY_IDS, _chk, c_dp
    Lh = slv.L

    def e2e_step():
        _chk(Lh.nekcem_b200_set_array(slv.h, ARRAY_IDS["hn"], hn_host.ctypes.data_as(c_dp), n3))
        _chk(Lh.nekcem_b200_set_array(slv.h, ARRAY_IDS["en"], en_host.ctypes.data_as(c_dp), n3))
        _chk(Lh.nekcem_b200_step(slv.h, 1))
        _chk(Lh.nekcem_b200_get_array(slv.h, ARRAY_IDS["hn"], hn_host.ctypes.data_as(c_dp), n3))
        _chk(Lh.nekcem_b200_get_array(slv.h, ARRAY_IDS["en"], en_host.ctypes.data_as(c_dp), n3))

    e2e_step()
    barrier()
    t0 = time.perf_counter()
    for _ in range(Ke):
        e2e_step()
    barrier()
    te = time.perf_counter() - t0
    tt = torch.tensor([te], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
    e2e_value = npts_global * 5.0 * Ke / float(tt.item()) / 1e9
    bytes_io = 2 * n3 * 8

    if rank == 0:
        peak, peak_src = measured_peak()
        bytes_stage = slv.algorithmic_bytes_per_stage()          # this rank's elements
        stage_ms = ms_max / (5.0 * K)
        achieved = bytes_stage / (stage_ms * 1e-3) / 1e9
        # exact redundancy the setup scan found in this mesh's geometry (same numbers, fewer
        # bytes): constant cofactors per element (-72 B/node there), hbm1 == ebm1 (-8 B/node)
        n_cm, shared = slv.geometry_info()
        n_el = slv.nelt
        actual_bytes = bytes_stage - 8.0 * nx1 ** 3 * (9.0 * n_cm + (n_el if shared else 0))
        traffic = None
        tpath = os.path.join(ROOT, "profiles", "traffic.json")
        if os.path.exists(tpath):
            try:
                traffic = None if strong and world > 1 else \
                    json.load(open(tpath)).get(
                        f"N{args.order}_E{E}" + ("_stream" if args.metrics == "stream" else ""))
            except Exception:
                traffic = None
        cpu = None
        if world == 1 and not args.no_cpu_baseline:
            ref = cpu_reference_rate(args.ref_elems, nx1, args.ref_steps, 1)
            if ref is not None:
                rate, threads, dtc, nptc = ref
                cpu = {"value": rate, "unit": UNIT, "cores": threads, "kind": "reference",
                       "sample": f"{threads} independent single-process replicas of the "
                                 "reference's own path (oracle/_ref: its Fortran translated to C "
                                 "+ its src/jl gs library; no MPI in this image), each a periodic "
                                 f"box of {args.ref_elems}^3 elements at N={args.order} ({nptc} "
                                 f"nodes in total), {args.ref_steps} steps, {dtc:.1f} s"}
            else:
                rate, threads, dtc, nptc = cpu_oracle_rate(args.cpu_elems, nx1, args.cpu_steps, 1)
                cpu = {"value": rate, "unit": UNIT, "cores": threads, "kind": "port",
                       "sample": f"periodic box {args.cpu_elems}^3 elements, N={args.order} "
                                 f"({nptc} nodes), {args.cpu_steps} steps, {dtc:.1f} s; oracle "
                                 "port (oracle/_ref did not travel with the tree)"}
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": K,
            "warmup": W, "ms_per_step": ms_max / K, "higher_is_better": True,
            "scaling": "strong" if strong else "weak",
            "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {
                "workload": f"synthetic 3D periodic box, {E}^3 hex elements "
                            f"{'in total' if strong else 'per GPU'} at N={args.order} "
                            f"(global {nel[0]}x{nel[1]}x{nel[2]}), 3dboxper initial condition, "
                            "upwind flux, LSRK(5,4)",
                "nodes_global": npts_global, "dof_unit": "grid node (6 field components)",
                "dt": dt, "partition": "reference pencil map (z-slabs), NCCL face exchange",
                "l2_flush": "inputs larger than L2 (one stage streams >> 126 MB)",
                "metrics_variant": (
                    f"{n_cm} of {n_el} elements have bitwise-constant cofactors (read once per "
                    f"element, SURVEY.md 8d 'affine-element shortcut'), hbm1==ebm1 "
                    f"{'shared' if shared else 'not shared'}; roofline.achieved counts the full "
                    "280+696/n B/node; run with --metrics stream for the general per-node path"
                    if args.metrics == "auto" else
                    "every geometry array streamed per node (general path, --metrics stream)"),
                "setup_s": round(t_setup, 1),
                "l2_error_vs_analytic": float(np.max(l2)), "linf_error_vs_analytic": float(np.max(linf)),
            },
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s",
                         "frac": achieved / peak, "traffic": traffic, "peak_source": peak_src,
                         "algorithmic_bytes_per_launch": bytes_stage,
                         "compulsory_bytes_per_launch_this_mesh": actual_bytes,
                         "frac_of_compulsory_this_mesh": actual_bytes / (stage_ms * 1e-3) / 1e9 / peak,
                         "kernel": "stage_kernel (one launch per RK stage per element list)",
                         "avg_launch_ms": stage_ms},
            "cpu_baseline": cpu,
            "clocks": clocks,
            "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": bytes_io,
                    "d2h_bytes_per_step": bytes_io, "steps": Ke,
                    "what": "per step: H2D(hn,en) from pinned host + nekcem_b200_step(1) + D2H(hn,en)"},
            "gpu_launches": int(launches),
        }
        print(json.dumps(line), file=_JSON_OUT, flush=True)
    slv.close()
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--elems", type=int, default=64, help="elements per direction per GPU")
    ap.add_argument("--order", type=int, default=7, help="polynomial order N (nx1 = N+1)")
    ap.add_argument("--e2e-steps", type=int, default=3)
    ap.add_argument("--cpu-elems", type=int, default=24)
    ap.add_argument("--cpu-steps", type=int, default=30)
    ap.add_argument("--ref-elems", type=int, default=12,
                    help="elements per direction of each replica of the reference CPU arm")
    ap.add_argument("--ref-steps", type=int, default=12)
    ap.add_argument("--ref-worker", default=None, help=argparse.SUPPRESS)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--metrics", default="auto", choices=["auto", "stream"],
                    help="auto: exploit exact redundancy of the geometry found at setup; "
                         "stream: read every geometry array per node (what a mesh with "
                         "round-off noise in its metrics gets)")
    ap.add_argument("--scaling", default="weak", choices=["weak", "strong"],
                    help="weak: --elems^3 per GPU (default); strong: --elems^3 in total")
    args = ap.parse_args()
    if args.ref_worker:
        e, n, k, w, d, i = args.ref_worker.split(",")
        _ref_worker(int(e), int(n), int(k), int(w), d, int(i))
        return
    if args.impl == "reference":
        run_reference(args)
    else:
        run_gpu(args)


if __name__ == "__main__":
    main()
