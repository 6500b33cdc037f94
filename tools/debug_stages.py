"""Stage-by-stage GPU-vs-oracle comparison of one time step (debug aid, not a test)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import navierstokes3d_b200 as ns
from oracle import oracle as O

nx = int(sys.argv[1]) if len(sys.argv) > 1 else 31
nsteps = int(sys.argv[2]) if len(sys.argv) > 2 else 2
variant = sys.argv[3] if len(sys.argv) > 3 else "M"
p = O.params_M(nx) if variant == "M" else O.params_G(nx)
f = O.initial_fields(p)
s = ns.setup_multi_gpu(nx) if variant == "M" else ns.setup_gpu(nx)
sim = ns.Simulation(s, ns.Context(0, ns.PARITY))
c, d = sim.ctx, sim.f
n = (s.nx, s.ny, s.nz)

def cmp(stage, names=("Pr", "dPrdtau", "Vx", "Vy", "Vz", "C", "divV", "Vx_o", "Vy_o", "Vz_o", "C_o", "txx", "txy")):
    ok = True
    for k in names:
        g = d[k].to_host()
        bad = np.argwhere(g != f[k])
        if len(bad):
            ok = False
            i = tuple(bad[0])
            print(f"  MISMATCH after {stage}: {k} {len(bad)} values, first {i}: gpu {g[i]!r} oracle {f[k][i]!r}; index ranges {bad.min(0)} .. {bad.max(0)}")
    print(f"{stage}: {'ok' if ok else 'DIFF'}")
    return ok

for step in range(nsteps):
    print("== step", step + 1)
    O.update_tau(p, f)
    c.call("ns3d_update_tau", d["txx"], d["tyy"], d["tzz"], d["txy"], d["txz"], d["tyz"], d["Vx"], d["Vy"], d["Vz"], s.mu, s.dx, s.dy, s.dz, *n)
    O.predict_V(p, f)
    c.call("ns3d_predict_V", d["Vx"], d["Vy"], d["Vz"], d["txx"], d["tyy"], d["tzz"], d["txy"], d["txz"], d["tyz"], s.rho, s.g, s.dt, s.dx, s.dy, s.dz, *n)
    cmp("predict_V")
    O.set_cylinder(p, f); sim.set_cylinder()
    O.update_divV(p, f)
    c.call("ns3d_update_divV", d["divV"], d["Vx"], d["Vy"], d["Vz"], s.dx, s.dy, s.dz, *n)
    cmp("divV")
    it_o, h_o = O.pt_solve(p, f)
    it_g, h_g = c.pt_solve(d["Pr"], d["dPrdtau"], d["divV"], s.pt_params())
    print("  iters", it_o, it_g)
    cmp("pt_solve")
    O.correct_V(p, f)
    c.call("ns3d_correct_V", d["Vx"], d["Vy"], d["Vz"], d["Pr"], s.dt, s.rho, s.dx, s.dy, s.dz, *n)
    cmp("correct_V")
    O.set_cylinder(p, f); sim.set_cylinder()
    cmp("set_cylinder")
    O.set_bc_Vel(p, f); sim.set_bc_Vel()
    cmp("set_bc_Vel")
    for a in ("Vx", "Vy", "Vz", "C"):
        f[a + "_o"][...] = f[a]
        c.copy(d[a + "_o"], d[a])
    cmp("copies")
    O.advect(p, f)
    c.call("ns3d_advect", d["Vx"], d["Vx_o"], d["Vy"], d["Vy_o"], d["Vz"], d["Vz_o"], d["C"], d["C_o"], s.dt, s.dx, s.dy, s.dz, *n)
    cmp("advect")
