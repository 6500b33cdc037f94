"""Stall-sample breakdown of one kernel from `ncu -i X.ncu-rep --page source --csv --print-source sass`:
totals per stall reason, and the SASS regions (between barriers / branches) that hold most samples.

    python tools/ncu_source_stalls.py /tmp/src.csv [kernel-index]
"""
import csv, sys
from collections import Counter
rows = list(csv.reader(open(sys.argv[1])))
# the file holds one block per kernel: a "Kernel Name" row, a header row, then instruction rows
blocks, cur = [], None
for r in rows:
    if r and r[0] == "Kernel Name":
        cur = {"name": r[1], "hdr": None, "rows": []}
        blocks.append(cur)
    elif cur is not None and cur["hdr"] is None:
        cur["hdr"] = r
    elif cur is not None:
        cur["rows"].append(r)
b = blocks[int(sys.argv[2]) if len(sys.argv) > 2 else 0]
h = {n: i for i, n in enumerate(b["hdr"])}
stall_cols = [n for n in b["hdr"] if n.startswith("stall_") and "Not Issued" not in n]
tot = Counter()
nsamp = 0
inst_total = 0
for r in b["rows"]:
    for c in stall_cols:
        tot[c] += int(r[h[c]] or 0)
    nsamp += int(r[h["# Samples"]] or 0)
    inst_total += int(r[h["Instructions Executed"]] or 0)
print(b["name"][:100])
print("samples", nsamp, "warp instructions", inst_total)
for c, v in tot.most_common():
    print(f"  {c:28s} {v:8d} {100.0 * v / max(nsamp, 1):6.2f} %")
# hot instructions
top = sorted(b["rows"], key=lambda r: -int(r[h["# Samples"]] or 0))[:40]
print("top instructions by samples:")
for r in top:
    why = max(stall_cols, key=lambda c: int(r[h[c]] or 0))
    print(f"  {r[h['Address']][-5:]} {int(r[h['# Samples']]):6d} exec {int(r[h['Instructions Executed']]):8d} {why:22s} {r[h['Source']].strip()[:90]}")
