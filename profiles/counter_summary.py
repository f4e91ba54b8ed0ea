"""Per-kernel-family summary of an `ncu --csv` log taken with a metric list (one row per launch and metric, or --page raw):
launches, share of the device time, average duration, DRAM bytes per launch, DRAM throughput %, tensor-pipe active %, issue-slot %.
Usage: python profiles/counter_summary.py log.csv > summary.txt"""
import csv
import re
import sys
from collections import defaultdict

rows = list(csv.reader(l for l in open(sys.argv[1]) if l.startswith('"')))
hdr = rows[0]
ix = {h: i for i, h in enumerate(hdr)}
per = defaultdict(lambda: defaultdict(list))
ids = {}
for r in rows[1:]:
    if len(r) < len(hdr):
        continue
    name = re.sub(r"\(.*", "", r[ix["Kernel Name"]]).replace("void ", "").replace("dp::<unnamed>::", "").replace("dp::", "")
    name = re.sub(r"<unnamed>::", "", name)
    try:
        v = float(r[ix["Metric Value"]].replace(",", ""))
    except ValueError:
        continue
    unit = r[ix["Metric Unit"]]
    m = r[ix["Metric Name"]]
    if m == "gpu__time_duration.sum":
        v = v / 1e3 if unit in ("ns", "nsecond") else (v * 1e3 if unit in ("ms", "msecond") else v)   # -> us
    if m.startswith("dram__bytes"):
        v = v * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(unit, 1)
    per[name][m].append(v)
tot = sum(sum(d["gpu__time_duration.sum"]) for d in per.values())
print(f"total device time {tot / 1e3:.3f} ms over {sum(len(d['gpu__time_duration.sum']) for d in per.values())} launches")
print(f"{'kernel':58s} {'n':>4s} {'share':>6s} {'avg us':>8s} {'DRAM MB':>9s} {'DRAM %':>7s} {'GB/s':>7s} {'tensor %':>8s} {'issue %':>7s} {'regs':>5s}")


def avg(d, k):
    return sum(d[k]) / len(d[k]) if d.get(k) else float("nan")


for name, d in sorted(per.items(), key=lambda kv: -sum(kv[1]["gpu__time_duration.sum"])):
    t = d["gpu__time_duration.sum"]
    by = avg(d, "dram__bytes_read.sum") + avg(d, "dram__bytes_write.sum")
    print(f"{name[:58]:58s} {len(t):4d} {100 * sum(t) / tot:5.1f}% {sum(t) / len(t):8.1f} {by / 1e6:9.2f} "
          f"{avg(d, 'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed'):7.1f} {by / (sum(t) / len(t)) / 1e3:7.0f} "
          f"{avg(d, 'sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active'):8.1f} "
          f"{avg(d, 'smsp__issue_active.avg.pct_of_peak_sustained_active'):7.1f} {avg(d, 'launch__registers_per_thread'):5.0f}")
