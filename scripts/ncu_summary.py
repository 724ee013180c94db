"""Summarise an ncu --set full report (developer tool): key metrics of the first kernel and the
share of warp-stall samples per kernel phase (phases are delimited by BAR.SYNC instructions).
Usage: python scripts/ncu_summary.py gpurun_out/prof_x.ncu-rep [> profiles/...txt]"""
import csv
import io
import subprocess
import sys

rep = sys.argv[1]
KEYS = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "lts__t_sector_hit_rate.pct", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
        "l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
        "l1tex__t_sector_hit_rate.pct",
        "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__warps_active.avg.pct_of_peak_sustained_active",
        "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "smsp__inst_executed.sum", "launch__registers_per_thread", "launch__block_size",
        "launch__grid_size", "launch__shared_mem_per_block_dynamic",
        "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem",
        "launch__occupancy_limit_warps", "sm__cycles_elapsed.avg",
        "smsp__inst_executed_op_local_ld.sum", "smsp__inst_executed_op_local_st.sum",
        "lts__t_sectors_srcunit_tex_op_read.sum", "lts__t_sectors_srcunit_tex_op_write.sum"]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units, vals = rows[0], rows[1], rows[2]
d = dict(zip(hdr, vals))
u = dict(zip(hdr, units))
print("Kernel Name =", d.get("Kernel Name"))
for k in hdr:
    if k in KEYS or "issue_stalled" in k and "per_issue_active" in k:
        print(f"{k} = {d[k]} {u[k]}")
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
lines = src.splitlines()
start = next(i for i, l in enumerate(lines) if l.startswith('"Address"'))
rd = list(csv.DictReader(io.StringIO("\n".join(lines[start:]))))
tot = sum(int(r["# Samples"]) for r in rd)
print(f"total samples {tot} ninstr {len(rd)}")
seg, segs = {"n": 0, "s": 0, "i": 0, "stalls": {}}, []
stall_cols = [c for c in rd[0].keys() if c.startswith("stall_") and "Not Issued" not in c]
for i, r in enumerate(rd):
    seg["s"] += int(r["# Samples"]); seg["i"] += int(r["Instructions Executed"]); seg["n"] += 1
    for c in stall_cols:
        seg["stalls"][c] = seg["stalls"].get(c, 0) + int(r[c] or 0)
    if "BAR.SYNC" in r["Source"] or i == len(rd) - 1:
        segs.append(seg); seg = {"n": 0, "s": 0, "i": 0, "stalls": {}}
for q, s in enumerate(segs):
    top = sorted(s["stalls"].items(), key=lambda kv: -kv[1])[:4]
    print(f"segment {q}: {s['n']} instr, samples {100.0*s['s']/tot:.1f} %, inst_exec {s['i']}, top stalls " +
          ", ".join(f"{k[6:]} {100.0*v/max(s['s'],1):.0f}%" for k, v in top))
if len(sys.argv) > 2:  # hottest instructions
    hot = sorted(rd, key=lambda r: -int(r["# Samples"]))[:int(sys.argv[2])]
    for r in hot:
        print(r["# Samples"], r["Source"].strip(), {c[6:]: r[c] for c in stall_cols if int(r[c] or 0) > 0.2 * int(r["# Samples"])})
