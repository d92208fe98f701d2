"""Summarise an ncu `--metrics gpu__time_duration.sum --csv` launch list: per-kernel totals of the LAST step
(from the last stem launch to the end: the earlier launches are the weight conversion and the warm-up pass)."""
import csv, collections, sys
rows = list(csv.reader(open(sys.argv[1])))
hi = next(i for i, r in enumerate(rows) if "Kernel Name" in r)
h = rows[hi]; ki = h.index("Kernel Name"); vi = h.index("Metric Value"); ui = h.index("Metric Unit")
recs = [(r[ki], float(r[vi].replace(",", "")) * ({"ns": 1e-3, "us": 1.0, "usecond": 1.0, "nsecond": 1e-3, "ms": 1e3, "msecond": 1e3}[r[ui]]))
        for r in rows[hi + 1:] if len(r) > vi and r[vi]]
stems = [i for i, (k, _) in enumerate(recs) if "stem" in k]
half = recs[stems[-1]:] if (len(sys.argv) < 3 and stems) else (recs[len(recs) // 2:] if len(sys.argv) < 3 else recs)
agg = collections.OrderedDict()
for k, us in half:
    name = k.split("(")[0].split("::")[-1]
    a = agg.setdefault(name, [0, 0.0]); a[0] += 1; a[1] += us
tot = sum(v[1] for v in agg.values())
print(f"{len(half)} launches, {tot:.1f} us total (serialised, cold-cache ncu replay times)")
for k, (n, t) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print(f"{k:44s} x{n:4d} {t:9.1f} us  {100*t/tot:5.1f} %  avg {t/n:8.1f}")
