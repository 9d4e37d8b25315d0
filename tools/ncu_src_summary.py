"""Summarise an `ncu --page source --csv` dump: stall-reason totals and the hottest SASS instructions."""
import csv
import sys

rows = list(csv.reader(open(sys.argv[1])))
hdr = rows[1]
idx = {h: i for i, h in enumerate(hdr)}
stall_cols = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
tot = {h: 0.0 for h in stall_cols}
items = []
alls = 0.0
for r in rows[2:]:
    try:
        v = float(r[idx["Warp Stall Sampling (All Samples)"]])
    except Exception:
        continue
    alls += v
    br = {}
    for h in stall_cols:
        try:
            x = float(r[idx[h]])
        except Exception:
            x = 0.0
        tot[h] += x
        if x:
            br[h[6:]] = x
    items.append((v, r[idx["Source"]].strip()[:90], br, r[idx["Instructions Executed"]]))
print("total samples", alls)
for h, v in sorted(tot.items(), key=lambda kv: -kv[1])[:10]:
    print(f"  {h:28s} {100 * v / max(alls, 1):6.2f}%")
top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
for v, s, br, ne in sorted(items, key=lambda t: -t[0])[:top]:
    b = " ".join(f"{k}:{int(x)}" for k, x in sorted(br.items(), key=lambda kv: -kv[1])[:3])
    print(f"{100 * v / alls:6.2f}%  exec={ne:>9s}  {s:90s} {b}")
