import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import navierstokes3d_b200 as ns
from oracle import oracle as O
nx, nt = 31, 2
p = O.params_M(nx)
f, iters_o, _ = O.run(p, nt)
f2, iters_o2, _ = O.run(p, nt)
print("oracle deterministic:", all((f[k] == f2[k]).all() for k in f), iters_o, iters_o2)
s = ns.setup_multi_gpu(nx)
for trial in range(3):
    sim = ns.Simulation(s, ns.Context(0, ns.PARITY))
    for _ in range(nt):
        sim.step()
    print("trial", trial, sim.iters)
    for k in ("Pr", "dPrdtau", "divV", "Vx_o", "Vy_o", "Vz_o", "C_o", "Vx", "Vy", "Vz", "C"):
        g = sim.host(k)
        bad = np.argwhere(g != f[k])
        if len(bad):
            i = tuple(bad[0])
            print(f"  {k}: {len(bad)} differ, first {i}: gpu {g[i]!r} oracle {f[k][i]!r} maxabs {np.abs(g-f[k]).max():.3e} range {bad.min(0)}..{bad.max(0)}")
    sim.ctx.close()
