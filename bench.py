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


def bind_to_gpu_numa(local: int):
    """Pin this rank to the cores of its GPU's NUMA node before the pinned host buffers are
    allocated (first touch places them there): the e2e copies then cross no inter-socket link."""
    try:
        import torch
        bus = torch.cuda.get_device_properties(local).pci_bus_id
        dom = torch.cuda.get_device_properties(local).pci_domain_id
        dev = torch.cuda.get_device_properties(local).pci_device_id
        path = f"/sys/bus/pci/devices/{dom:04x}:{bus:02x}:{dev:02x}.0/numa_node"
        node = int(open(path).read().strip())
        if node < 0:
            return None
        cpus = []
        for part in open(f"/sys/devices/system/node/node{node}/cpulist").read().strip().split(","):
            lo, _, hi = part.partition("-")
            cpus += list(range(int(lo), int(hi or lo) + 1))
        cpus = sorted(set(cpus) & os.sched_getaffinity(0))
        if cpus:
            os.sched_setaffinity(0, cpus)
            return node
    except Exception:
        pass
    return None


def bytes_per_node(nx1: int) -> float:
    """algorithmic bytes per node and stage of the general path, SURVEY.md 8d: 280 + 696/n"""
    return 280.0 + 696.0 / nx1


def run_gpu(args):
    import torch
    import torch.distributed as dist

    from nekcem_b200 import MaxwellB200, comm_unique_id
    from nekcem_b200.boxcase import BoxCase, BoxCase2D, gll

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world != args.gpus:
        if world == 1 and args.gpus > 1:
            raise SystemExit("launch with torch.distributed.run --nproc-per-node N for --gpus N")
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device (B200); there is no CPU fallback")
    torch.cuda.set_device(local)
    numa = bind_to_gpu_numa(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(x: float) -> float:
        t = torch.tensor([x], dtype=torch.float64, device="cuda")
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    peak, peak_src = measured_peak()
    transports = set()

    def make(nel, order, nranks=world, myrank=rank, general=True, **variant):
        """solver for a periodic box of nel elements at order N, partitioned over nranks"""
        nx1 = order + 1
        t0 = time.perf_counter()
        case = BoxCase(nel, nx1, rank=myrank, nranks=nranks, length=2 * math.pi, **variant)
        slv = MaxwellB200(3, nx1, case.nelt, device=local, rank=myrank, nranks=nranks,
                          ifpml=bool(variant.get("pml")))
        slv.cem_maxwell_init(case.lazy(), free_after_upload=True)
        if nranks > 1:
            uid = torch.zeros(128, dtype=torch.uint8, device="cuda")
            if rank == 0:
                uid.copy_(torch.frombuffer(bytearray(comm_unique_id()), dtype=torch.uint8))
            dist.broadcast(uid, 0)
            slv.comm_init(bytes(uid.cpu().numpy().tobytes()))
        if general:
            slv.set_option("const_metrics", 0)
        if args.transport == "nccl":
            slv.set_option("p2p", 0)
        slv.setup()
        transports.add(slv.transport())
        # CFL-limited dt as in the synthetic .rea (param(12)=+0.1 -> dt = 0.1*dxmin)
        z, _ = gll(nx1)
        dxmin = 0.5 * min(case.h) * 0.5 * float(np.min(z[2:] - z[:-2])) if nx1 > 2 else min(case.h)
        dt = 0.1 * dxmin
        slv.set_time(0.0, dt)
        return case, slv, dt, time.perf_counter() - t0

    def timed(slv, warm, steps):
        """device time of `steps` time steps (CUDA events on the compute stream), max over ranks"""
        slv.step(warm)
        barrier()
        slv.cem_maxwell_op_rk(steps, sync=False)
        slv.synchronize()
        barrier()
        ms, launches = slv.last_step_ms()
        return max_over_ranks(ms), int(launches)

    def measure(nel, order, warm, steps, general=True, **variant):
        """one extra configuration: rate, ms/step and the roofline fraction of its stage kernel"""
        case, slv, dt, ts = make(nel, order, general=general, **variant)
        ms, launches = timed(slv, warm, steps)
        npts_global = int(np.prod(nel)) * (order + 1) ** 3
        if variant:   # no closed-form solution: the fields must stay finite and bounded
            zero = np.zeros(3 * case.npts)
            ssum, smax = slv.error_sums(zero, zero)
            del zero
        else:
            shn, sen = case.fields(slv.time)
            ssum, smax = slv.error_sums(shn, sen)
            del shn, sen
        red = torch.tensor(ssum, dtype=torch.float64, device="cuda")
        if world > 1:
            dist.all_reduce(red, op=dist.ReduceOp.SUM)
        l2 = float(np.sqrt(red.cpu().numpy().max() / case.volume_global))
        npml = int(case.array("pmlptr").size)
        slv.close()
        stage_ms = ms / (5.0 * steps)
        # + 240 B per node of a PML element
        bpn = bytes_per_node(order + 1) + 240.0 * npml / max(case.nelt, 1)
        # the slowest rank's launch carries npts_global/world nodes (equal slabs)
        ach = bpn * (npts_global / world) / (stage_ms * 1e-3) / 1e9
        return {"value": npts_global * 5.0 * steps / (ms * 1e-3) / 1e9, "unit": UNIT,
                "ms_per_step": ms / steps, "steps": steps, "nodes_global": npts_global,
                "elements_global": list(nel), "order": order,
                "roofline_frac": ach / peak, "achieved_GBps_per_gpu": ach,
                "bytes_per_node_stage": bpn,
                ("l2_norm_of_fields" if variant else "l2_error_vs_analytic"): l2,
                "pml_elements_per_rank": npml, "variant": {k: str(v) for k, v in variant.items()},
                "gpu_launches": launches, "setup_s": round(ts, 1)}

    def measure_2d(E2, order, warm, steps):
        """the 2D TE path (stage2d_kernel; tests/drude, tests/2dboxper are 2D) on a periodic box of
        E2 x E2 elements.  Algorithmic bytes: 144 B/node + 84 B/face point = 144 + 336/n."""
        nx1 = order + 1
        t0 = time.perf_counter()
        case = BoxCase2D((E2, E2), nx1, imode=1)
        slv = MaxwellB200(2, nx1, case.nelt, imode=1, device=local)
        slv.cem_maxwell_init(case.arrays())
        slv.setup()
        z, _ = gll(nx1)
        slv.set_time(0.0, 0.1 * 0.5 * min(case.h) * 0.5 * float(np.min(z[2:] - z[:-2])))
        ts = time.perf_counter() - t0
        ms, launches = timed(slv, warm, steps)
        shn, sen = case.fields(slv.time)
        ssum, _ = slv.error_sums(shn, sen)
        l2 = float(np.sqrt(ssum.max() / case.volume_global))
        slv.close()
        stage_ms = ms / (5.0 * steps)
        bpn = 144.0 + 336.0 / nx1
        ach = bpn * case.npts / (stage_ms * 1e-3) / 1e9
        return {"value": case.npts * 5.0 * steps / (ms * 1e-3) / 1e9, "unit": UNIT,
                "ms_per_step": ms / steps, "steps": steps, "nodes_global": case.npts,
                "elements_global": [E2, E2], "order": order, "roofline_frac": ach / peak,
                "achieved_GBps_per_gpu": ach, "bytes_per_node_stage": bpn,
                "l2_error_vs_analytic": l2, "gpu_launches": launches, "setup_s": round(ts, 1),
                "kernel": "stage2d_kernel (TE)"}

    E, order = args.elems, args.order
    nx1 = order + 1
    strong = args.scaling == "strong"
    # weak scaling (default): one E^3 slab per GPU; strong: the global E^3 box split over the GPUs
    nel = (E, E, E) if strong else (E, E, E * world)
    general = args.metrics == "stream"
    case, slv, dt, t_setup = make(nel, order, general=general)
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
    ms_max = max_over_ranks(ms)
    value = npts_global * 5.0 * K / (ms_max * 1e-3) / 1e9
    n_cm, shared = slv.geometry_info()

    # the same mesh through the constant-metric instantiation (exact redundancy of THIS mesh's
    # geometry, never found in metrics that came out of the reference's glmapm1): labelled extra
    cm_variant = None
    if general and not args.no_extras:
        slv.set_option("const_metrics", 1)
        ms_cm, _ = timed(slv, 3, K)
        n_cm1, shared1 = slv.geometry_info()
        cm_bytes = (bytes_per_node(nx1) - (72.0 if n_cm1 == slv.nelt else 0.0)
                    - (8.0 if shared1 else 0.0)) * case.npts
        cm_variant = {"value": npts_global * 5.0 * K / (ms_cm * 1e-3) / 1e9, "unit": UNIT,
                      "ms_per_step": ms_cm / K,
                      "elements_with_constant_cofactors": int(n_cm1), "hbm1_eq_ebm1": bool(shared1),
                      "bytes_per_launch_this_variant": cm_bytes,
                      "roofline_frac_of_bytes_moved": cm_bytes / (ms_cm / (5.0 * K) * 1e-3) / 1e9 / peak,
                      "note": "same results bit for bit; reachable only when the uploaded cofactors "
                              "are bitwise constant per element"}
        slv.set_option("const_metrics", 0)
        slv.step(1)

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
    # every step takes its input fields from pinned host memory and hands the advanced fields
    # back to pinned host memory (the `!$ACC UPDATE DEVICE/HOST(hn,en)` seams of the reference,
    # drude.usr:94, 3dboxper.usr:199) through nekcem_b200_step_streamed: the upload of input
    # k+1, the five stages of input k and the download of result k-1 overlap (three streams,
    # PCIe full duplex); consecutive inputs are independent states, as a stream of them would be
    Ke = K if args.e2e_steps <= 0 else max(1, min(args.e2e_steps, K))
    n3 = 3 * case.npts
    e2e_buffers = "separate pinned input and output buffers"
    try:
        pins = [torch.empty(n3, dtype=torch.float64, pin_memory=True) for _ in range(4)]
    except RuntimeError:
        # not enough pinnable host memory for 4 x 3*npts doubles per rank: results come back into
        # the input buffers (the upload of input k+1 then reads what the download of result k-1
        # is writing; same bytes over PCIe, the values of later inputs are undefined)
        pins = [torch.empty(n3, dtype=torch.float64, pin_memory=True) for _ in range(2)]
        pins = pins + pins
        e2e_buffers = "outputs written into the input buffers (host memory limit)"
    hn_in, en_in, hn_out, en_out = (p.numpy() for p in pins)
    hn_in[:] = slv.hn
    en_in[:] = slv.en

    def e2e_step():
        slv.step_streamed(hn_in, en_in, hn_out, en_out)

    for _ in range(3):          # fills the pipeline (input -> compute -> result) and warms up
        e2e_step()
    barrier()
    t0 = time.perf_counter()
    for _ in range(Ke):
        e2e_step()
    barrier()
    te = max_over_ranks(time.perf_counter() - t0)
    e2e_value = npts_global * 5.0 * Ke / te / 1e9
    bytes_io = 2 * n3 * 8
    # the streamed result is one time step applied to the uploaded state: check it
    slv.step_streamed(None, None, hn_out, en_out)     # drain: result of the last input
    e2e_ok = bool(np.isfinite(hn_out[:1024]).all() and np.abs(hn_out).max() > 0)
    del pins, hn_in, en_in, hn_out, en_out
    slv.close()

    # ---- further configurations of BASELINE.json configs[4] in the same line -----------------
    extra = {}
    if not args.no_extras:
        Kx = max(3, min(K, 5))
        if not (order == 15 and E == 32 and not strong):
            extra["weak_n15_e32_per_gpu"] = measure((32, 32, 32 * world), 15, 3, Kx)
        if world == 1:
            # the auxiliary paths at benchmark size (BASELINE.json configs[2], [3]): every element a
            # PML element (tests/3dboxpml), two materials + PML layers (tests/3ddielectric)
            extra["aux_all_pml_n7_e40"] = measure((40, 40, 40), 7, 3, Kx, pml="all")
            extra["aux_two_materials_pml_layers_n7_e40"] = measure(
                (40, 40, 40), 7, 3, Kx, pml="layers", eps_upper=4.0)
            extra["aux_2d_te_n7_e512"] = measure_2d(512, 7, 3, Kx)
        if world > 1:
            extra["strong_n7_e64_total"] = measure((64, 64, 64), 7, 3, Kx)
        if world >= 4:
            extra["strong_n15_e64_total"] = measure((64, 64, 64), 15, 3, Kx)
        if world > 1:
            extra["parity_vs_single_domain"] = parity_vs_single_domain(
                torch, dist, make, rank, world, barrier)

    if rank == 0:
        bpn = bytes_per_node(nx1)
        bytes_stage = bpn * case.npts                    # this rank's elements, general path
        stage_ms = ms_max / (5.0 * K)
        if general:
            moved = bytes_stage
        else:  # --metrics auto on a mesh with exact redundancy: count what the variant must move
            moved = bytes_stage - 8.0 * nx1 ** 3 * (9.0 * n_cm + (slv.nelt if shared else 0))
        achieved = moved / (stage_ms * 1e-3) / 1e9
        traffic, traffic_src = None, None
        tpath = os.path.join(ROOT, "profiles", "traffic.json")
        if os.path.exists(tpath) and not (strong and world > 1):
            try:
                ent = json.load(open(tpath)).get(f"N{order}_E{E}_" + ("general" if general else "auto"))
                if ent:
                    traffic, traffic_src = ent["dram_bytes_per_launch"], ent["source"]
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
                                 f"box of {args.ref_elems}^3 elements at N={order} ({nptc} "
                                 f"nodes in total), {args.ref_steps} steps, {dtc:.1f} s"}
            else:
                rate, threads, dtc, nptc = cpu_oracle_rate(args.cpu_elems, nx1, args.cpu_steps, 1)
                cpu = {"value": rate, "unit": UNIT, "cores": threads, "kind": "port",
                       "sample": f"periodic box {args.cpu_elems}^3 elements, N={order} "
                                 f"({nptc} nodes), {args.cpu_steps} steps, {dtc:.1f} s; oracle "
                                 "port (oracle/_ref did not travel with the tree)"}
        xmirror = 2.0 * 6 * 8 * 2 * nx1 * nx1 * slv.nelt     # write + read of the x-face mirror
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": K,
            "warmup": W, "ms_per_step": ms_max / K, "higher_is_better": True,
            "scaling": "strong" if strong else "weak",
            "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {
                "workload": f"synthetic 3D periodic box, {E}^3 hex elements "
                            f"{'in total' if strong else 'per GPU'} at N={order} "
                            f"(global {nel[0]}x{nel[1]}x{nel[2]}), 3dboxper initial condition, "
                            "upwind flux, LSRK(5,4)",
                "nodes_global": npts_global, "dof_unit": "grid node (6 field components)",
                "dt": dt, "partition": "reference pencil map (z-slabs)",
                "face_exchange": sorted(transports),
                "l2_flush": "inputs larger than L2 (one stage streams >> 126 MB)",
                "metrics_variant": (
                    "general path: every geometry array streamed per node (what a mesh whose "
                    "metrics came out of the reference's glmapm1 gets); the constant-metric "
                    "instantiation on this mesh is reported under variants" if general else
                    f"{n_cm} of {slv.nelt} elements have bitwise-constant cofactors, read once per "
                    f"element; hbm1==ebm1 {'shared' if shared else 'not shared'}; roofline counts "
                    "the bytes this variant must move"),
                "setup_s": round(t_setup, 1), "numa_node": numa,
                "l2_error_vs_analytic": float(np.max(l2)), "linf_error_vs_analytic": float(np.max(linf)),
            },
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s",
                         "frac": achieved / peak, "traffic": traffic, "traffic_source": traffic_src,
                         "peak_source": peak_src,
                         "algorithmic_bytes_per_launch": moved,
                         "bytes_per_node_stage": moved / case.npts,
                         "overhead_bytes_per_launch_not_credited": xmirror,
                         "overhead_what": "x-face mirror of the fields (compact copy of the +-x "
                                          "traces, written by the epilogue, read by the x "
                                          "neighbours instead of a stride-n gather)",
                         "kernel": "pipe_kernel (nx1 5, 7..10) / slab_kernel: one launch per RK stage "
                                   "per element list",
                         "avg_launch_ms": stage_ms},
            "variants": {"constant_metrics": cm_variant},
            "extra": extra,
            "cpu_baseline": cpu,
            "clocks": clocks,
            "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": bytes_io,
                    "d2h_bytes_per_step": bytes_io, "steps": Ke, "result_checked": e2e_ok,
                    "host_buffers": e2e_buffers,
                    "what": "per step: H2D(hn,en) of the next input from pinned host memory + one "
                            "time step + D2H(hn,en) of the previous result to pinned host memory, "
                            "through nekcem_b200_step_streamed (three streams; inputs of "
                            "consecutive steps are independent states)"},
            "gpu_launches": int(launches),
        }
        print(json.dumps(line), file=_JSON_OUT, flush=True)
    if world > 1:
        dist.destroy_process_group()


def parity_vs_single_domain(torch, dist, make, rank, world, barrier):
    """The same global mesh (8 x 8 x 8*world elements, N=7, general path) advanced K steps by the
    `world` ranks with the NCCL face exchange and by rank 0 alone: per-node arithmetic is
    identical, so the fields must agree bit for bit (what gs_op_fields guarantees the reference,
    src/cem_maxwell.F:962, src/jl/gs.c:471-512)."""
    K = 3
    nel = (8, 8, 8 * world)
    case, slv, dt, _ = make(nel, 7)
    slv.step(K)
    loc = torch.from_numpy(np.concatenate([slv.hn, slv.en])).cuda()
    ids = torch.from_numpy(case.lglel.copy()).cuda()
    slv.close()
    allf = [torch.empty_like(loc) for _ in range(world)]
    alli = [torch.empty_like(ids) for _ in range(world)]
    dist.all_gather(allf, loc)
    dist.all_gather(alli, ids)
    out = None
    if rank == 0:
        nxyz = case.nxyz
        nelg = int(np.prod(nel))
        glob = np.empty((6, nelg * nxyz))
        for f, i in zip(allf, alli):
            f = f.cpu().numpy().reshape(6, -1, nxyz)
            glob.reshape(6, nelg, nxyz)[:, i.cpu().numpy(), :] = f
        c1, s1, _, _ = make(nel, 7, nranks=1, myrank=0)
        s1.set_time(0.0, dt)
        s1.step(K)
        one = np.concatenate([s1.hn, s1.en]).reshape(6, -1)
        s1.close()
        diff = float(np.max(np.abs(glob - one)))
        den = float(np.sqrt(np.sum(one * one)))
        out = {"bitwise_equal": bool(np.array_equal(glob, one)), "max_abs_diff": diff,
               "rel_l2": float(np.sqrt(np.sum((glob - one) ** 2)) / den),
               "mesh": f"{nel[0]}x{nel[1]}x{nel[2]} elements, N=7, {K} steps, {world} ranks vs 1"}
    barrier()
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--elems", type=int, default=64, help="elements per direction per GPU")
    ap.add_argument("--order", type=int, default=7, help="polynomial order N (nx1 = N+1)")
    ap.add_argument("--e2e-steps", type=int, default=0, help="0: as many as --steps")
    ap.add_argument("--no-extras", action="store_true",
                    help="skip the extra configurations (N=15, strong scaling, multi-rank parity)")
    ap.add_argument("--cpu-elems", type=int, default=24)
    ap.add_argument("--cpu-steps", type=int, default=30)
    ap.add_argument("--ref-elems", type=int, default=12,
                    help="elements per direction of each replica of the reference CPU arm")
    ap.add_argument("--ref-steps", type=int, default=12)
    ap.add_argument("--ref-worker", default=None, help=argparse.SUPPRESS)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--metrics", default="stream", choices=["auto", "stream"],
                    help="stream (default): read every geometry array per node -- the general "
                         "path, what a mesh with round-off noise in its metrics (any mesh out of "
                         "the reference's glmapm1) gets; auto: exploit exact redundancy of the "
                         "geometry found at setup")
    ap.add_argument("--transport", default="p2p", choices=["p2p", "nccl"],
                    help="inter-GPU face exchange: stores into the peers' halo buffers over NVLink "
                         "(default) or grouped ncclSend/ncclRecv")
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
