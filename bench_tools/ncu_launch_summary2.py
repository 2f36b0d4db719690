"""Per-kernel totals of an `ncu --metrics gpu__time_duration.sum --csv` launch list (every launch of the run)."""
import collections, csv, gzip, json, sys
path = sys.argv[1]
op = gzip.open if path.endswith(".gz") else open
rows = list(csv.reader(op(path, "rt")))
hdr = None; agg = collections.OrderedDict(); n = 0
for r in rows:
    if r and r[0] == "ID": hdr = r; continue
    if not hdr or len(r) != len(hdr): continue
    name = r[hdr.index("Kernel Name")].split("(")[0]
    v = float(r[hdr.index("Metric Value")].replace(",", "")); u = r[hdr.index("Metric Unit")]
    us = v / 1000 if u == "ns" else (v * 1000 if u == "ms" else v)
    a = agg.setdefault(name, [0, 0.0]); a[0] += 1; a[1] += us; n += 1
tot = sum(a[1] for a in agg.values())
own = sum(a[1] for k, a in agg.items() if "grasp::" in k or k.startswith("grasp") or "void grasp" in k)
out = {"launches": n, "total_us": tot, "own_kernel_share_of_time": own / tot if tot else None,
       "kernels": [{"name": k, "launches": a[0], "total_us": round(a[1], 1), "share": round(a[1] / tot, 4)}
                   for k, a in sorted(agg.items(), key=lambda kv: -kv[1][1])]}
print(json.dumps(out, indent=1))
