"""Stall-reason totals (warp-sampling, all samples) of one kernel of an .ncu-rep: usage ncu_stalls.py rep [kernel_index]"""
import collections, csv, io, subprocess, sys
rep, k = sys.argv[1], int(sys.argv[2]) if len(sys.argv) > 2 else 0
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--kernel-id", f":::{k + 1}"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(src)))
h = rows[1]
cols = [c for c in h if c.startswith("stall_") and "Not Issued" not in c]
seen, agg, tot_i = set(), collections.Counter(), 0
for r in rows[2:]:
    if len(r) < len(h) or not r[0].startswith("0x") or r[0] in seen:
        continue
    seen.add(r[0])
    tot_i += int(r[h.index("Instructions Executed")] or 0)
    for c in cols:
        v = r[h.index(c)]
        if v:
            agg[c] += int(v)
tot = sum(agg.values())
print(rows[0][1][:70], "| warp-instructions", tot_i, "| samples", tot)
for c, v in agg.most_common(10):
    print(f"  {c:28s} {v:7d} {100 * v / tot:5.1f}%")
