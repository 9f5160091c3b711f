"""Summarise an .ncu-rep: one line per kernel + top stall instructions. usage: ncu_summary.py rep [kernel_index_for_source]"""
import csv, io, subprocess, sys
rep = sys.argv[1]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units, data = rows[0], rows[1], rows[2:]
def col(d, name):
    return d[hdr.index(name)] if name in hdr else "-"
def mb(d, name):          # byte metrics come in varying units (Kbyte / Mbyte / Gbyte): always print MB
    if name not in hdr:
        return "-"
    i = hdr.index(name)
    scale = {"byte": 1e-6, "Kbyte": 1e-3, "Mbyte": 1.0, "Gbyte": 1e3}.get(units[i], 1.0)
    return "%.3f" % (float(d[i].replace(",", "")) * scale)
for i, d in enumerate(data):
    print(i, col(d, "Kernel Name")[:44],
          "%s=%s" % (units[hdr.index("gpu__time_duration.sum")], col(d, "gpu__time_duration.sum")[:7]),
          "tensor%%=%s" % col(d, "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active")[:5],
          "inst=%s" % col(d, "smsp__inst_executed.sum")[:10],
          "lat/inst=%s" % col(d, "smsp__average_warp_latency_per_inst_issued.ratio")[:5],
          "warps%%=%s" % col(d, "sm__warps_active.avg.pct_of_peak_sustained_active")[:5],
          "dramR=%s W=%s MB" % (mb(d, "dram__bytes_read.sum"), mb(d, "dram__bytes_write.sum")),
          "lts%%=%s" % col(d, "lts__throughput.avg.pct_of_peak_sustained_elapsed")[:5],
          "regs=%s" % col(d, "launch__registers_per_thread"))
if len(sys.argv) > 2:
    k = int(sys.argv[2])
    src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--kernel-id", f":::{k+1}"], capture_output=True, text=True).stdout
    rows = [r for r in csv.reader(io.StringIO(src))]
    h = rows[1]
    d = [r for r in rows[2:] if len(r) > 5 and r[0].startswith("0x")]
    si, ii = h.index("Warp Stall Sampling (All Samples)"), h.index("Instructions Executed")
    seen, out = set(), []
    for r in d:
        if r[0] in seen: continue
        seen.add(r[0]); out.append(r)
    tot = sum(int(r[si] or 0) for r in out)
    print("total samples", tot)
    for r in sorted(out, key=lambda r: -int(r[si] or 0))[:int(sys.argv[3]) if len(sys.argv) > 3 else 25]:
        print(r[si].rjust(7), r[ii].rjust(9), r[0][-5:], r[1][:110])
