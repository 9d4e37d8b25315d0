"""Condense an .ncu-rep (`ncu --set full`) into the handful of numbers DESIGN.md / bench.py cite.
   python tools/ncu_summary.py report.ncu-rep out.json"""
import csv
import io
import json
import subprocess
import sys

KEYS = {
    "gpu__time_duration.sum": "duration",
    "dram__bytes_read.sum": "dram_read",
    "dram__bytes_write.sum": "dram_write",
    "dram__throughput.avg.pct_of_peak_sustained_elapsed": "dram_pct",
    "lts__throughput.avg.pct_of_peak_sustained_elapsed": "l2_pct",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed": "sm_pct",
    "sm__warps_active.avg.pct_of_peak_sustained_active": "warps_active_pct",
    "smsp__issue_active.avg.pct_of_peak_sustained_active": "issue_active_pct",
    "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active": "fp64_pipe_pct",
    "sm__pipe_tensor_subpipe_dmma_cycles_active.avg.pct_of_peak_sustained_active": "dmma_pipe_pct",
    "sm__mem_tensor_cycles_active.avg.pct_of_peak_sustained_active": "tensor_mem_pct",
    "launch__registers_per_thread": "registers",
    "launch__block_size": "block",
    "launch__grid_size": "grid",
    "launch__cluster_size": "cluster",
    "launch__shared_mem_per_block_dynamic": "dyn_smem",
    "smsp__inst_executed.sum": "warp_instructions",
}
UNIT = {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1.0, "ms": 1e-3, "us": 1e-6, "ns": 1e-9, "s": 1.0,
        "msecond": 1e-3, "usecond": 1e-6, "nsecond": 1e-9, "second": 1.0}

raw = subprocess.run(["ncu", "-i", sys.argv[1], "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units = rows[0], rows[1]
out = []
for r in rows[2:]:
    d = dict(zip(hdr, r))
    e = {"kernel": d["Kernel Name"][:120]}
    for k, name in KEYS.items():
        if k in d and d[k] != "":
            u = units[hdr.index(k)]
            try:
                v = float(d[k].replace(",", ""))
            except ValueError:
                continue
            e[name] = v * UNIT.get(u, 1.0) if name in ("duration", "dram_read", "dram_write", "dyn_smem") else v
    if "dram_read" in e and "dram_write" in e:
        e["dram_traffic_bytes"] = e["dram_read"] + e["dram_write"]
        if e.get("duration"):
            e["dram_gbs"] = e["dram_traffic_bytes"] / e["duration"] / 1e9
    out.append(e)
json.dump(out, open(sys.argv[2], "w"), indent=1)
for e in out:
    print(json.dumps(e))
