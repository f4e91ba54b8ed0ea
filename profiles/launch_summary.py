"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list by kernel + grid."""
import collections
import csv
import re
import sys

rows = list(csv.reader(open(sys.argv[1])))
hdr = None
agg = collections.OrderedDict()
tot = 0.0
n = 0
for r in rows:
    if hdr is None:
        if "Kernel Name" in r:
            hdr = r
        continue
    d = dict(zip(hdr, r))
    if d.get("Metric Name") != "gpu__time_duration.sum":
        continue
    v = float(d["Metric Value"].replace(",", ""))
    v = v / 1e3 if d["Metric Unit"] == "ns" else (v * 1e3 if d["Metric Unit"] == "ms" else v)
    key = (re.sub(r"\(.*", "", d["Kernel Name"]).replace("void dp::<unnamed>::", "").replace("dp::<unnamed>::", ""), d["Grid Size"])
    a = agg.setdefault(key, [0, 0.0])
    a[0] += 1
    a[1] += v
    tot += v
    n += 1
print(f"launches {n}  total {tot:.0f} us")
for k, (c, t) in sorted(agg.items(), key=lambda kv: -kv[1][1])[: int(sys.argv[2]) if len(sys.argv) > 2 else 25]:
    print(f"{t:9.0f} us {100 * t / tot:5.1f}%  x{c:3d}  avg {t / c:8.1f}  {k[0][:60]} grid={k[1]}")
