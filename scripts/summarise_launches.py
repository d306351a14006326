"""Aggregate an ncu launch list (--metrics gpu__time_duration.sum --csv) per kernel.
usage: summarise_launches.py raw.csv out.csv"""
import csv, sys, collections
rows = [r for r in csv.reader(open(sys.argv[1])) if r]
hdr_i = next(i for i, r in enumerate(rows) if "Kernel Name" in r)
hdr = rows[hdr_i]
ik, ig, iv, iu = hdr.index("Kernel Name"), hdr.index("Grid Size"), hdr.index("Metric Value"), hdr.index("Metric Unit")
agg = collections.OrderedDict()
scale = {"ns": 1e-6, "us": 1e-3, "usecond": 1e-3, "ms": 1.0, "msecond": 1.0, "nsecond": 1e-6, "s": 1e3, "second": 1e3}
for r in rows[hdr_i + 1:]:
    if len(r) <= iv:
        continue
    key = (r[ik], r[ig])
    ms = float(r[iv].replace(",", "")) * scale.get(r[iu], 1e-6)
    a = agg.setdefault(key, [0, 0.0])
    a[0] += 1
    a[1] += ms
tot = sum(a[1] for a in agg.values()) or 1.0
with open(sys.argv[2], "w", newline="") as f:
    w = csv.writer(f)
    w.writerow(["kernel", "grid", "launches", "total_ms", "share"])
    for (k, g), (n, ms) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        w.writerow([k, g, n, f"{ms:.3f}", f"{ms / tot:.4f}"])
