"""A short run of the fused PT iteration for ncu: profile_pt.py GRID MODE ZCHUNK _ name=value,name=value (12 iterations)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import navierstokes3d_b200 as ns
g = sys.argv[1] if len(sys.argv) > 1 else "255x153x153"
mode = sys.argv[2] if len(sys.argv) > 2 else "FAST"
zc = int(sys.argv[3]) if len(sys.argv) > 3 else 0
nx, ny, nz = map(int, g.split("x"))
s = ns.setup_gpu(nx, ny=ny, nz=nz)
ctx = ns.Context(0, getattr(ns, mode))
ctx.set_option("graphs", 0)
if len(sys.argv) > 5:   # further library options: name=value,name=value
    for kv in sys.argv[5].split(","):
        k, v = kv.split("=")
        ctx.set_option(k, int(v))
rng = np.random.default_rng(0)
Pr = ctx.from_host(np.asfortranarray(rng.uniform(-1, 1, size=(nx, ny, nz))))
dP = ctx.zeros(nx - 2, ny - 2, nz - 2)
dv = ctx.from_host(np.asfortranarray(rng.uniform(-1e-3, 1e-3, size=(nx, ny, nz))))
ctx.pt_iterate(Pr, dP, dv, s.pt_params(zc), int(sys.argv[6]) if len(sys.argv) > 6 else 12)
ctx.sync()
print("done", ctx.launch_count)
ctx.close()
