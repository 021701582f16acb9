"""Summarises an `ncu --metrics gpu__time_duration.sum --csv` launch list into per-kernel totals and shares."""
import csv
import collections
import sys

path = sys.argv[1]
rows = []
with open(path, newline="") as f:
    lines = [l for l in f if not l.startswith("==")]
rd = csv.DictReader(lines)
tot = collections.defaultdict(lambda: [0, 0.0])
for r in rd:
    if r.get("Metric Name") != "gpu__time_duration.sum":
        continue
    name = r["Kernel Name"].split("(")[0]
    v = float(r["Metric Value"].replace(",", ""))
    unit = r.get("Metric Unit", "ns")
    ns = v * {"ns": 1.0, "us": 1e3, "ms": 1e6, "nsecond": 1.0, "usecond": 1e3, "msecond": 1e6}.get(unit, 1.0)
    tot[name][0] += 1
    tot[name][1] += ns
total = sum(v[1] for v in tot.values())
print(f"# {path}: {sum(v[0] for v in tot.values())} launches, {total / 1e6:.3f} ms summed kernel time")
print(f"{'kernel':60s} {'launches':>9s} {'total ms':>10s} {'avg us':>9s} {'share':>7s}")
for name, (cnt, ns) in sorted(tot.items(), key=lambda kv: -kv[1][1]):
    print(f"{name[:60]:60s} {cnt:9d} {ns / 1e6:10.3f} {ns / cnt / 1e3:9.2f} {ns / total:7.1%}")
