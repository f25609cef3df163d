"""Aggregates an `ncu --metrics gpu__time_duration.sum --csv` launch list by kernel name: total us, launches."""
import csv, collections, sys
rows = list(csv.reader(open(sys.argv[1])))
hdr, agg = None, collections.defaultdict(lambda: [0, 0.0])
for r in rows:
    if "Kernel Name" in r:
        hdr = r
        continue
    if hdr is None or len(r) < len(hdr):
        continue
    d = dict(zip(hdr, r))
    if d.get("Metric Name") != "gpu__time_duration.sum":
        continue
    v = float(d["Metric Value"].replace(",", ""))
    v = {"ns": v / 1e3, "us": v, "ms": v * 1e3, "s": v * 1e6}.get(d["Metric Unit"], v)
    k = d["Kernel Name"][:70]
    agg[k][0] += 1
    agg[k][1] += v
tot = sum(t for _, t in agg.values())
for k, (c, t) in sorted(agg.items(), key=lambda x: -x[1][1])[:int(sys.argv[2]) if len(sys.argv) > 2 else 30]:
    print(f"{t:10.1f} us {100 * t / tot:5.1f}% {c:6d}  {k}")
