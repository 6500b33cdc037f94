"""ns3d_step vs oracle after every step (debug aid)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import navierstokes3d_b200 as ns
from oracle import oracle as O

nx = int(sys.argv[1]) if len(sys.argv) > 1 else 31
nsteps = int(sys.argv[2]) if len(sys.argv) > 2 else 2
p = O.params_M(nx)
f = O.initial_fields(p)
s = ns.setup_multi_gpu(nx)
sim = ns.Simulation(s, ns.Context(0, ns.PARITY))
for k in ("Pr", "Vx", "Vy", "Vz", "C"):
    g = sim.host(k)
    print("init", k, "ok" if (g == f[k]).all() else "DIFF")
for step in range(nsteps):
    it_o, _ = O.step(p, f)
    it_g, _ = sim.step()
    print("== step", step + 1, it_o, it_g)
    for k in ("Pr", "dPrdtau", "divV", "txx", "txy", "Vx_o", "Vy_o", "Vz_o", "C_o", "Vx", "Vy", "Vz", "C"):
        g = sim.host(k)
        bad = np.argwhere(g != f[k])
        if len(bad):
            i = tuple(bad[0])
            print(f"  {k}: {len(bad)} differ, first {i}: gpu {g[i]!r} oracle {f[k][i]!r} range {bad.min(0)}..{bad.max(0)}")
        else:
            print(f"  {k}: ok")
