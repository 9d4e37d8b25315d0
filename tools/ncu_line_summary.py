"""Aggregate `ncu --page source --csv --print-source cuda,sass` per CUDA source line:
   python tools/ncu_line_summary.py dump.csv [top]"""
import csv
import sys
from collections import defaultdict

rows = csv.reader(open(sys.argv[1]))
top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
cur_file = None
hdr = None
acc = defaultdict(lambda: [0.0, 0.0, ""])  # (file, line) -> [samples, instr, text]
cur_line = None
stalls = defaultdict(float)
for r in rows:
    if not r:
        continue
    if r[0] == "File Path":
        cur_file = r[1].split("/")[-1]
        continue
    if r[0] == "Function Name":
        continue
    if r[0] == "Line No":
        hdr = r
        i_addr = hdr.index("Address")
        i_samp = hdr.index("Warp Stall Sampling (All Samples)")
        i_inst = hdr.index("Instructions Executed")
        stall_idx = [(i, h) for i, h in enumerate(hdr) if h.startswith("stall_") and "Not Issued" not in h]
        continue
    if hdr is None:
        continue
    if r[0] != "":
        cur_line = (cur_file, int(r[0]))
        acc[cur_line][2] = r[1].strip()[:100]
    if len(r) > i_addr and r[i_addr] != "" and cur_line is not None:
        try:
            s = float(r[i_samp]); n = float(r[i_inst])
        except ValueError:
            continue
        acc[cur_line][0] += s
        acc[cur_line][1] += n
        for i, h in stall_idx:
            try:
                stalls[h] += float(r[i])
            except ValueError:
                pass
tot_s = sum(v[0] for v in acc.values()) or 1.0
tot_i = sum(v[1] for v in acc.values()) or 1.0
print(f"total samples {tot_s:.0f}, warp instructions {tot_i:.0f}")
for h, v in sorted(stalls.items(), key=lambda kv: -kv[1])[:8]:
    print(f"  {h:24s} {100 * v / tot_s:6.2f}%")
byfile = defaultdict(lambda: [0.0, 0.0])
for (f, l), v in acc.items():
    byfile[f][0] += v[0]; byfile[f][1] += v[1]
for f, v in sorted(byfile.items(), key=lambda kv: -kv[1][0]):
    print(f"  file {f:24s} samples {100 * v[0] / tot_s:6.2f}%  instr {100 * v[1] / tot_i:6.2f}%")
for (f, l), v in sorted(acc.items(), key=lambda kv: -kv[1][0])[:top]:
    print(f"{100 * v[0] / tot_s:6.2f}% smp {100 * v[1] / tot_i:6.2f}% ins  {f}:{l:<4d} {v[2]}")
