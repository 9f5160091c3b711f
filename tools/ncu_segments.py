"""Split a one-kernel .ncu-rep source page at its block barriers: samples and instructions per segment + top stalls."""
import csv, collections, io, subprocess, sys
rep = sys.argv[1]
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--kernel-id", ":::1"], capture_output=True, text=True).stdout
rows = [r for r in csv.reader(io.StringIO(src))]
h = rows[1]
si, ii = h.index("Warp Stall Sampling (All Samples)"), h.index("Instructions Executed")
cols = [c for c in h if c.startswith('stall_') and 'Not Issued' not in c]
d = [r for r in rows[2:] if len(r) > 5 and r[0].startswith('0x')]
seen, out = set(), []
for r in d:
    if r[0] in seen: continue
    seen.add(r[0]); out.append(r)
tot = sum(int(r[si] or 0) for r in out); ins = sum(int(r[ii] or 0) for r in out)
print("samples", tot, "warp-instr", ins, "n", len(out))
acc_s = acc_i = 0; agg = collections.Counter(); start = 0
for i, r in enumerate(out):
    acc_s += int(r[si] or 0); acc_i += int(r[ii] or 0)
    for c in cols:
        v = r[h.index(c)]
        if v: agg[c] += int(v)
    if 'BAR.SYNC' in r[1] or i == len(out) - 1:
        print(f"seg {start:5d}-{i:5d} samples {acc_s:5d} ({100*acc_s/tot:4.1f}%) instr {acc_i/1e6:6.2f}M  BARx{r[ii]}  {agg.most_common(4)}")
        acc_s = acc_i = 0; agg = collections.Counter(); start = i + 1
n = int(sys.argv[2]) if len(sys.argv) > 2 else 20
for r in sorted(out, key=lambda r: -int(r[si] or 0))[:n]:
    print(r[si].rjust(6), r[ii].rjust(9), str(out.index(r)).rjust(5), r[1][:100])
