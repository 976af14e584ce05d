"""Summarises an `ncu --metrics gpu__time_duration.sum --csv` launch list by kernel name.
python tools/summarize_launches.py launches.csv [end_kernel]
With end_kernel (e.g. cfg_step_kernel: the last launch of a denoising step) the list is cut into
steps at every launch of that kernel, the first (set-up + first step) segment is dropped and the
table is the per-step average over the remaining complete steps."""
import collections, csv, re, sys
path = sys.argv[1]
end_kernel = sys.argv[2] if len(sys.argv) > 2 else None
with open(path, newline="") as f:
    lines = [l for l in f if not l.startswith("==")]
launches = []
for r in csv.DictReader(lines):
    if r.get("Metric Name") != "gpu__time_duration.sum":
        continue
    name = re.sub(r"\(.*", "", r["Kernel Name"])
    name = re.sub(r"^void ", "", name)
    v = float(r["Metric Value"].replace(",", ""))
    unit = r["Metric Unit"]
    us = v / 1e3 if unit in ("nsecond", "ns") else v if unit in ("usecond", "us") else v * 1e3
    launches.append((name, us))
steps = 1
if end_kernel:
    ends = [i for i, (n, _) in enumerate(launches) if end_kernel in n]
    segs = [launches[a + 1:b + 1] for a, b in zip(ends[:-1], ends[1:])]
    launches = [x for s in segs for x in s]
    steps = len(segs)
tot = collections.defaultdict(lambda: [0.0, 0])
for name, us in launches:
    tot[name][0] += us
    tot[name][1] += 1
total = sum(v[0] for v in tot.values())
print(f"# {path}: {len(launches)} launches in {steps} step(s), {total / 1e3 / steps:.3f} ms of kernel time per step")
print("# per-launch times under ncu are cold-cache and serialised: compare SHARES, not absolutes")
print(f"{'kernel':70s} {'launches/step':>13s} {'ms/step':>9s} {'share':>7s} {'avg us':>8s}")
for k, (us, n) in sorted(tot.items(), key=lambda kv: -kv[1][0]):
    print(f"{k[:70]:70s} {n / steps:13.1f} {us / 1e3 / steps:9.3f} {100 * us / total:6.1f}% {us / n:8.1f}")
