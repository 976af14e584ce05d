"""Summarises an `ncu --metrics gpu__time_duration.sum --csv` launch list by kernel name."""
import collections, csv, re, sys
path = sys.argv[1]
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 1
rows = []
with open(path, newline="") as f:
    lines = [l for l in f if not l.startswith("==")]
rd = csv.DictReader(lines)
tot = collections.defaultdict(lambda: [0.0, 0])
for r in rd:
    if r.get("Metric Name") != "gpu__time_duration.sum":
        continue
    name = re.sub(r"\(.*", "", r["Kernel Name"])
    name = re.sub(r"^void ", "", name)
    v = float(r["Metric Value"].replace(",", ""))
    unit = r["Metric Unit"]
    us = v / 1e3 if unit in ("nsecond", "ns") else v if unit in ("usecond", "us") else v * 1e3
    tot[name][0] += us
    tot[name][1] += 1
total = sum(v[0] for v in tot.values())
print(f"# {path}: {sum(v[1] for v in tot.values())} launches, {total / 1e3 / steps:.3f} ms of kernel time per step ({steps} steps captured)")
print(f"{'kernel':70s} {'launches/step':>13s} {'ms/step':>9s} {'share':>7s} {'avg us':>8s}")
for k, (us, n) in sorted(tot.items(), key=lambda kv: -kv[1][0]):
    print(f"{k[:70]:70s} {n / steps:13.1f} {us / 1e3 / steps:9.3f} {100 * us / total:6.1f}% {us / n:8.1f}")
