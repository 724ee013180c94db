"""profiles/traffic.json from an ncu --csv log holding dram__bytes_read/write per launch
(developer tool).  usage: python scripts/make_traffic.py <log.csv> <key, e.g. N7_E64>"""
import csv, json, os, sys
log, key = sys.argv[1], sys.argv[2]
rows = [r for r in csv.reader(l for l in open(log) if l.startswith('"'))]
hdr = rows[0]
ik, im, iv = hdr.index("Kernel Name"), hdr.index("Metric Name"), hdr.index("Metric Value")
iid = hdr.index("ID")
per = {}
for r in rows[1:]:
    if "slab_kernel" not in r[ik] and "stage2d" not in r[ik]:
        continue
    d = per.setdefault(r[iid], {})
    d[r[im]] = float(r[iv].replace(",", ""))
tot = [d["dram__bytes_read.sum"] + d["dram__bytes_write.sum"] for d in per.values()
       if "dram__bytes_read.sum" in d and "dram__bytes_write.sum" in d]
tot = tot[len(tot) // 2:]  # skip warm-up launches
path = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "profiles", "traffic.json")
cur = json.load(open(path)) if os.path.exists(path) else {}
cur[key] = sum(tot) / len(tot)
cur[key + "_source"] = f"ncu dram__bytes_read.sum + dram__bytes_write.sum, mean of {len(tot)} launches ({os.path.basename(log)})"
json.dump(cur, open(path, "w"), indent=1)
print(key, cur[key])
