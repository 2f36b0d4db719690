"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list (optionally gzipped) per kernel."""
import collections, csv, gzip, io, json, sys
path, out_path, command, note = sys.argv[1], sys.argv[2], sys.argv[3], sys.argv[4]
text = (gzip.open(path, "rt") if path.endswith(".gz") else open(path)).read()
text = text[text.index('"ID","Process ID"'):]
rows = list(csv.reader(io.StringIO(text)))
hdr = rows[0]
kn, mv, mu = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
agg = collections.defaultdict(lambda: [0, 0.0])
unit = None
for r in rows[1:]:
    if len(r) <= mv:
        continue
    try:
        v = float(r[mv].replace(",", ""))
    except ValueError:
        continue
    unit = r[mu]
    name = r[kn].split("(")[0][:80]
    if "distribution_elementwise" in name:      # model initialisation, outside the bench's timed region
        continue
    a = agg[name]; a[0] += 1; a[1] += v
tot = sum(a[1] for a in agg.values())
out = {"command": command, "note": note, "unit": unit, "total": tot, "kernels": []}
for k, a in sorted(agg.items(), key=lambda kv: -kv[1][1])[:40]:
    out["kernels"].append({"kernel": k, "launches": a[0], "total": round(a[1], 1), "share": round(a[1] / tot, 4)})
json.dump(out, open(out_path, "w"), indent=1)
print(unit, tot, len(rows))
for k in out["kernels"][:16]:
    print(k)
