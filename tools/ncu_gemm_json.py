"""profiles/r02_ncu_update_gemms.json from the `ncu --set full` capture of one minibatch's 8 grouped-GEMM launches
(tools/ncu_capture.sh): per-launch duration, DRAM bytes, tensor-pipe activity; bench.py reads `traffic_bytes_per_launch`.
usage: ncu_gemm_json.py gpurun_out/ncu_update_gemms.ncu-rep > profiles/r02_ncu_update_gemms.json"""
import csv, io, json, subprocess, sys

NAMES = ["fwd L1 (actor 705->512 | critic 219->768 | lin-vel 705->128)", "fwd L2 (512->256 | 768->256 | 128->128)",
         "fwd L3 (256->128 x2)", "dW3 (split-K, atomic)", "dX2 (DELU + column sums)", "dW2", "dX1", "dW1"]
raw = subprocess.run(["ncu", "-i", sys.argv[1], "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units, data = rows[0], rows[1], rows[2:]


def val(d, name):
    i = hdr.index(name)
    v, u = float(d[i].replace(",", "")), units[i]
    scale = {"Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "byte": 1.0, "ms": 1e3, "us": 1.0, "usecond": 1.0, "ns": 1e-3, "nsecond": 1e-3,
             "msecond": 1e3}.get(u, 1.0)
    return v * scale


out = []
for i, d in enumerate(data):
    out.append({"launch": NAMES[i] if i < len(NAMES) else str(i), "kernel": d[hdr.index("Kernel Name")],
                "duration_us": val(d, "gpu__time_duration.sum"), "dram_read_bytes": val(d, "dram__bytes_read.sum"),
                "dram_write_bytes": val(d, "dram__bytes_write.sum"),
                "tensor_pipe_active_pct": val(d, "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active"),
                "registers": int(val(d, "launch__registers_per_thread"))})
r, w = sum(o["dram_read_bytes"] for o in out), sum(o["dram_write_bytes"] for o in out)
print(json.dumps({"source": "ncu --set full --clock-control none, one minibatch (61,440 rows) of PPO.update at 4096 envs x 60: "
                            "tools/ncu_capture.sh (tools/prof_iter.py, -k regex:gemm_kernel --launch-skip 63 -c 8; cold-cache, serialised replays)",
                  "launches": out, "traffic_bytes_per_launch": (r + w) / len(out), "dram_read_bytes_total": r,
                  "dram_write_bytes_total": w}, indent=1))
