"""Sum an `ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --csv` log over the launches of
one kernel and write it in the format of tools/ncu_summary.py (bench.py reads `dram_traffic_bytes`):
   python tools/ncu_metrics_sum.py log.csv kernel_substring scans out.json"""
import csv
import json
import sys

path, sub, scans, out = sys.argv[1], sys.argv[2], int(sys.argv[3]), sys.argv[4]
tot, n, name = {}, 0, None
for r in csv.reader(l for l in open(path) if not l.startswith("==")):
    if len(r) > 5 and r[0] != "ID" and sub in r[4]:
        name = r[4][:120]
        tot[r[-3]] = tot.get(r[-3], 0.0) + float(r[-1].replace(",", ""))
        n += 1
launches = n // 3
e = {"kernel": name, "launches_per_scan": launches // scans,
     "duration": tot.get("gpu__time_duration.sum", 0.0) * 1e-9 / scans,
     "dram_read": tot.get("dram__bytes_read.sum", 0.0) / scans, "dram_write": tot.get("dram__bytes_write.sum", 0.0) / scans,
     "note": "summed over the launches of one scan (one launch per eigen-tile group, digit planes pinned in the L2 set-aside); "
             "ncu --metrics pass, --clock-control none"}
e["dram_traffic_bytes"] = e["dram_read"] + e["dram_write"]
e["dram_gbs"] = e["dram_traffic_bytes"] / max(e["duration"], 1e-12) / 1e9
json.dump([e], open(out, "w"), indent=1)
print(json.dumps(e))
