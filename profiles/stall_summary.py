"""Summarise an `ncu --page source --csv` dump: stall-reason totals, samples per opcode, hottest SASS lines."""
import collections
import csv
import sys

rows = list(csv.reader(open(sys.argv[1])))
h = next(i for i, r in enumerate(rows) if "Address" in r and "Source" in r)
hdr = rows[h]
idx = {k: i for i, k in enumerate(hdr)}
stalls = [k for k in hdr if k.startswith("stall_") and "Not Issued" not in k]
tot = collections.Counter()
total = 0
data = []
for r in rows[h + 1:]:
    if len(r) < len(hdr) or not r[idx["# Samples"]].strip().isdigit():
        continue
    n = int(r[idx["# Samples"]])
    total += n
    st = {s: int(r[idx[s]] or 0) for s in stalls}
    tot.update(st)
    data.append((n, r[idx["Address"]], r[idx["Source"]][:100], st))
print("total samples", total)
print({k: v for k, v in tot.most_common() if v > 0})
byop = collections.Counter()
for n, a, s, st in data:
    tok = s.split()
    op = tok[1] if tok and tok[0].startswith("@") and len(tok) > 1 else (tok[0] if tok else "")
    byop[op.split(".")[0]] += n
print(byop.most_common(14))
for n, a, s, st in sorted(data, key=lambda d: -d[0])[: int(sys.argv[2]) if len(sys.argv) > 2 else 20]:
    print(n, a[-5:], s, sorted(st.items(), key=lambda kv: -kv[1])[:2])
