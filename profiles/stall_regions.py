"""Summarise an ncu source-page CSV (ncu -i x.ncu-rep --page source --csv --kernel-name regex:K) by code region between marker
instructions and list the hottest instructions with their top stall reasons.  Usage: python profiles/stall_regions.py file.csv [top]"""
import csv
import re
import sys

rows = list(csv.reader(open(sys.argv[1])))
top = int(sys.argv[2]) if len(sys.argv) > 2 else 30
hdr = rows[1]
idx = {h: i for i, h in enumerate(hdr)}
stalls = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
seen, data = set(), []
for r in rows[2:]:
    if len(r) < len(hdr) or r[idx["Address"]] in seen:
        continue
    seen.add(r[idx["Address"]])
    try:
        ns = int(r[idx["# Samples"]])
    except ValueError:
        continue
    data.append((r[idx["Address"]][-5:], r[idx["Source"]], ns, int(r[idx["Instructions Executed"]] or 0), r))
tot = sum(d[2] for d in data)
print("kernel", rows[0][1][:90], "samples", tot, "warp-instructions", sum(d[3] for d in data))
agg = {s: 0 for s in stalls}
for d in data:
    for s in stalls:
        agg[s] += int(d[4][idx[s]] or 0)
print("stall reasons:", ", ".join(f"{s[6:]} {100 * v / tot:.1f}%" for s, v in sorted(agg.items(), key=lambda x: -x[1])[:8]))
print("-- regions (samples since the previous marker, ending at the marker) --")
acc = acci = 0
for a, src, ns, ie, r in data:
    acc += ns
    acci += ie
    if re.search(r"SYNCS\.(PHASECHK|ARRIVE)|LDTM|USETMAXREG|MEMBAR|LDGDEPBAR|EXIT|STTM|BAR\.SYNC|UTCBAR", src) or (
        "UTCHMMA" in src and acc > 0.002 * tot
    ):
        if acc > 0.003 * tot:
            print(f"{a} {acc:7d} {100 * acc / tot:5.1f}%  inst {acci:10d} | {src.strip()[:90]}")
        acc = acci = 0
print("-- hottest instructions --")
for a, src, ns, ie, r in sorted(data, key=lambda x: -x[2])[:top]:
    t2 = sorted(((int(r[idx[s]] or 0), s[6:]) for s in stalls), reverse=True)[:2]
    print(f"{ns:6d} {100 * ns / tot:5.1f}% {a} {src.strip()[:70]:70s} {t2}")
