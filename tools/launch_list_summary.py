"""Per-kernel summary of an `ncu --metrics gpu__time_duration.sum --csv --log-file X` launch list.

    python tools/launch_list_summary.py gpurun_out/launches.csv profiles/rNN_bench_launch_list_summary.csv "<command>"
"""
import csv, re, sys
from collections import defaultdict

src, out = sys.argv[1], sys.argv[2]
cmd = sys.argv[3] if len(sys.argv) > 3 else ""
rows = [l for l in open(src, newline="") if l.startswith('"')]
rd = csv.reader(rows)
hdr = next(rd)
ki, vi, ui = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
acc = defaultdict(lambda: [0, 0.0])
for r in rd:
    name = re.sub(r"\(.*$", "", r[ki]).replace("void ", "").replace("<unnamed>::", "")
    v = float(r[vi].replace(",", ""))
    v *= {"ns": 1e-3, "us": 1.0, "ms": 1e3, "s": 1e6}.get(r[ui], 1.0)
    acc[name][0] += 1
    acc[name][1] += v
tot = sum(v[1] for v in acc.values())
with open(out, "w") as fh:
    fh.write(f"# {cmd}\n# per-launch times under ncu are cold-cache and serialised: compare SHARES\n")
    fh.write("kernel,launches,total_us,share_pct,avg_us\n")
    for name, (n, t) in sorted(acc.items(), key=lambda kv: -kv[1][1]):
        fh.write(f'"{name}",{n},{t:.1f},{100 * t / tot:.2f},{t / n:.2f}\n')
print(open(out).read())
