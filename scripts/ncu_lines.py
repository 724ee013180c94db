"""Top source lines by warp-stall samples from an ncu report (developer tool)."""
import csv, io, subprocess, sys
rep, top = sys.argv[1], int(sys.argv[2]) if len(sys.argv) > 2 else 30
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"],
                     capture_output=True, text=True).stdout
rows = []
cur = None
hdr = None
for r in csv.reader(io.StringIO(out)):
    if len(r) >= 2 and r[0] == "File Path":
        cur = r[1].split("/")[-1]; hdr = None; continue
    if len(r) >= 2 and r[0] == "Function Name": continue
    if r and r[0] == "Line No": hdr = r; continue
    if hdr and len(r) == len(hdr) and r[0] != "":
        d = {}
        for k, v in zip(hdr, r):
            d.setdefault(k, v)
        try: s = int(d["# Samples"])
        except Exception: continue
        rows.append((s, cur, d["Line No"], d["Source"].strip()[:90], d.get("stall_long_sb", ""), d.get("stall_barrier", ""),
                     int(d.get("L1 Wavefronts Shared") or 0), int(d.get("L1 Wavefronts Shared Ideal") or 0), int(d.get("Instructions Executed") or 0)))
tot = sum(r[0] for r in rows)
key = (lambda x: -x[6]) if len(sys.argv) > 3 and sys.argv[3] == "smem" else (lambda x: -x[8]) if len(sys.argv) > 3 and sys.argv[3] == "inst" else (lambda x: -x[0])
wtot = sum(r[6] for r in rows); itot = sum(r[8] for r in rows)
print(f"total smem wavefronts {wtot} ideal {sum(r[7] for r in rows)}  inst {itot}")
for s, f, ln, src, lsb, bar, wf, wfi, ins in sorted(rows, key=key)[:top]:
    print(f"{100.0*s/tot:5.1f}%  {f}:{ln} lsb={lsb} bar={bar} wf={100.0*wf/max(wtot,1):.1f}% (x{wf/max(wfi,1):.2f}) inst={100.0*ins/itot:.1f}%  {src}")
