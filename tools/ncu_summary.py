"""Compact text summary of an `ncu --set full` report (per kernel launch): duration, clocks, pipe
utilisation, DRAM traffic, registers, plus the warp-stall mix of the source page.
python tools/ncu_summary.py report.ncu-rep > profiles/<name>.txt"""
import csv, io, subprocess, sys, collections
rep = sys.argv[1]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units = rows[0], rows[1]
KEYS = ["Kernel Name", "Grid Size", "Block Size", "gpu__time_duration.sum", "sm__cycles_elapsed.max",
        "sm__cycles_elapsed.max.per_second", "sm__cycles_active.avg", "launch__registers_per_thread",
        "launch__shared_mem_per_block_dynamic", "launch__occupancy_limit_registers",
        "sm__warps_active.avg.pct_of_peak_sustained_active",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum",
        "dram__bytes_read.sum", "dram__bytes_write.sum", "dram__bytes_read.sum.per_second",
        "dram__bytes_write.sum.per_second", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "lts__t_sector_hit_rate.pct", "l1tex__data_bank_conflicts_pipe_lsu.sum"]
print(f"# {rep}  (ncu --set full --clock-control none; clocks and durations are under the profiler)")
for r in rows[2:]:
    d = dict(zip(hdr, r))
    for k in KEYS:
        if k in d:
            print(f"{k:70s} {d[k]:>20s} {units[hdr.index(k)]}")
    print()
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(src)))
try:
    hi = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
except StopIteration:
    sys.exit(0)
hdr = rows[hi]; ci = {h: i for i, h in enumerate(hdr)}
stall_cols = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
tot = collections.Counter(); n = 0; per = []
for idx, r in enumerate(rows[hi + 1:]):
    if len(r) < len(hdr): continue
    s = int(r[ci["# Samples"]] or 0); n += s
    st = {h: int(r[ci[h]] or 0) for h in stall_cols}
    for h, v in st.items(): tot[h] += v
    per.append((idx, r[ci["Source"]].strip(), s, int(r[ci["Instructions Executed"]] or 0), st))
print(f"warp-stall samples (first profiled launch): {n}")
for h, v in tot.most_common():
    if v: print(f"  {h:26s} {v:8d} {100.0 * v / n:5.1f}%")
print("hottest SASS lines:")
for idx, srcl, s, ex, st in sorted(per, key=lambda t: -t[2])[:20]:
    main = ", ".join(f"{k[6:]}={v}" for k, v in sorted(st.items(), key=lambda kv: -kv[1])[:2] if v)
    print(f"  {idx:5d} {100.0 * s / n:4.1f}% exec={ex:8d} {srcl[:58]:58s} {main}")
