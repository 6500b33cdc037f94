"""Records the DRAM traffic per launch of the fused PT kernel from an `ncu --set full` capture under the key bench.py
looks up (workload : mode : options : hash of the kernel sources), so that `roofline.traffic` is only ever reported
for the kernel source it was measured on.

    python tools/update_traffic.py gpurun_out/r2ncu/ptv_x_1.ncu-rep B FAST [--passes=N] [name=value ...]

--passes=N: the capture holds EVERY ptv_kernel launch of N passes over the fields (z-bands: several launches per pass).
"""
import csv, io, json, os, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402

rep, workload, mode = sys.argv[1:4]
opts = [a for a in sys.argv[4:] if not a.startswith("--passes=")]
passes = next((int(a.split("=")[1]) for a in sys.argv[4:] if a.startswith("--passes=")), 0)   # z-bands: launches per pass vary
SCALE = {"byte": 1, "kbyte": 1e3, "mbyte": 1e6, "gbyte": 1e9}
vals = []
if rep.endswith(".csv"):
    # `ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum --csv --log-file X.csv`: one row per launch and metric
    rows = list(csv.reader(l for l in open(rep, newline="") if l.startswith('"')))
    ix = {h: i for i, h in enumerate(rows[0])}
    per_launch = {}
    for r in rows[1:]:
        if "ptv_kernel" not in r[ix["Kernel Name"]] or not r[ix["Metric Name"]].startswith("dram__bytes"):
            continue
        v = float(r[ix["Metric Value"]].replace(",", "")) * SCALE[r[ix["Metric Unit"]].lower()]
        per_launch[r[ix["ID"]]] = per_launch.get(r[ix["ID"]], 0.0) + v
    vals = list(per_launch.values())
else:
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units, data = rows[0], rows[1], rows[2:]
    ix = {h: i for i, h in enumerate(hdr)}
    for r in data:
        if "ptv_kernel" not in r[ix["Kernel Name"]]:
            continue
        tot = 0.0
        for k in ("dram__bytes_read.sum", "dram__bytes_write.sum"):
            tot += float(r[ix[k]]) * SCALE[units[ix[k]].lower()]
        vals.append(tot)
traffic = sum(vals) / (passes if passes else len(vals))   # per pass over the fields (= per launch without z-bands)
path = os.path.join(ROOT, "profiles", "traffic.json")
d = json.load(open(path))
key = bench.kernel_key(workload, mode, opts)
d[key] = traffic
d["_doc"] = ("dram__bytes_read.sum + dram__bytes_write.sum per launch of ptv_kernel from `ncu --set full --clock-control none` "
             "(bytes). Keys: <bench workload>:<mode>:<options>:<sha1 of ns3d_ptv_kernels.cuh + ns3d_ptv.cu, 12 hex digits> "
             "(tools/update_traffic.py); bench.py reports null when the sources have changed since the capture.")
json.dump(d, open(path, "w"), indent=1)
print(key, traffic, f"from {len(vals)} launches of", rep)
