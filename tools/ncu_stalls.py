"""Summarise per-instruction warp-stall samples from `ncu --page source --csv` output:
python tools/ncu_stalls.py report.ncu-rep [top_n]"""
import csv, subprocess, sys, io, collections
rep = sys.argv[1]; top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
txt = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(txt)))
hi = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
hdr = rows[hi]; body = rows[hi + 1:]
ci = {h: i for i, h in enumerate(hdr)}
stall_cols = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
tot = collections.Counter(); total_samples = 0
per = []
for n, r in enumerate(body):
    if len(r) < len(hdr): continue
    smp = int(r[ci["# Samples"]] or 0); total_samples += smp
    st = {h: int(r[ci[h]] or 0) for h in stall_cols}
    for h, v in st.items(): tot[h] += v
    per.append((n, r[ci["Source"]].strip(), smp, int(r[ci["Instructions Executed"]] or 0), st))
print("total samples", total_samples)
for h, v in tot.most_common(): 
    if v: print(f"  {h:28s} {v:8d}  {100.0 * v / total_samples:5.1f}%")
print("\ntop instructions by samples:")
for n, src, smp, ex, st in sorted(per, key=lambda t: -t[2])[:top]:
    main = ", ".join(f"{k[6:]}={v}" for k, v in sorted(st.items(), key=lambda kv: -kv[1])[:3] if v)
    print(f"{n:5d} {smp:7d} ({100.0 * smp / total_samples:4.1f}%) exec={ex:8d}  {src[:60]:60s} {main}")
if len(sys.argv) > 3:
    lo, hi2 = int(sys.argv[3]), int(sys.argv[4])
    print("\nrange:")
    for n, src, smp, ex, st in per[lo:hi2]:
        main = ", ".join(f"{k[6:]}={v}" for k, v in sorted(st.items(), key=lambda kv: -kv[1])[:3] if v)
        print(f"{n:5d} {smp:7d} exec={ex:8d}  {src[:60]:60s} {main}")
