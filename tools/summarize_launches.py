"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list: per kernel count, total, average, share.
Usage: python tools/summarize_launches.py profiles/<tag>_launches.csv > profiles/<tag>_launches_summary.txt
Per-launch times under ncu are cold-cache and serialised: compare SHARES, not absolutes.  Input generation
(k_mul_batch) and torch's own fills/copies are listed but excluded from the shares."""
import csv, re, sys
from collections import OrderedDict

rows = []
with open(sys.argv[1], newline="") as f:
    lines = [l for l in f if l.startswith('"')]
for r in csv.DictReader(lines):
    if r.get("Metric Name") == "gpu__time_duration.sum":
        rows.append((r["Kernel Name"], float(r["Metric Value"].replace(",", ""))))


def short(name):
    name = re.sub(r"^void ", "", name)
    name = re.sub(r"\(.*$", "", name)
    return name[:66]


agg = OrderedDict()
for k, ns in rows:
    a = agg.setdefault(short(k), [0, 0.0])
    a[0] += 1
    a[1] += ns
excluded = lambda k: k.startswith("at::") or "k_mul_batch" in k
total = sum(v[1] for k, v in agg.items() if not excluded(k))
print("# ncu --metrics gpu__time_duration.sum --clock-control none  (%s)" % sys.argv[1])
print("# per-launch times are cold-cache and serialised: compare SHARES (input generation k_mul_batch and torch fills excluded)")
print("%-66s %6s %12s %10s %7s" % ("kernel", "count", "total_ms", "avg_us", "share"))
for k, (c, ns) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    share = "-" if excluded(k) else "%.1f%%" % (100 * ns / total)
    print("%-66s %6d %12.3f %10.1f %7s" % (k, c, ns / 1e6, ns / c / 1e3, share))
acc = sum(v[1] for k, v in agg.items() if "pairtree_round" in k or "accumulate_slices" in k or "k_pt_" in k)
print("# bucket-accumulate pass (pair-tree rounds + XYZZ slices) share of one MSM under ncu: %.1f%%" % (100 * acc / total))
