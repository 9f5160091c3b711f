"""Summarise an ncu `--metrics gpu__time_duration.sum --csv` launch list: per-kernel totals (cold-cache, serialised)."""
import collections, csv, sys
rows = [r for r in csv.reader(open(sys.argv[1])) if len(r) > 14 and r[0].isdigit()]
agg = collections.OrderedDict()
for r in rows:
    k = (r[4][:64], r[8], r[7])
    a = agg.setdefault(k, [0, 0.0]); a[0] += 1; a[1] += float(r[14]) / 1000
tot = sum(v[1] for v in agg.values())
print(f"{len(rows)} launches, {tot:.1f} us summed")
for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1])[:int(sys.argv[2]) if len(sys.argv) > 2 else 30]:
    print("%-66s grid %-12s blk %-12s n=%4d total %8.1f us avg %7.1f  %4.1f%%" % (k[0], k[1], k[2], v[0], v[1], v[1] / v[0], 100 * v[1] / tot))
