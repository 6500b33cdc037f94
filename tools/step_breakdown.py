"""Where does a time step go? (config B, FASTEST) -- wall-clock with syncs around the three parts."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import navierstokes3d_b200 as ns
s = ns.setup_gpu(255)
ctx = ns.Context(0, ns.FASTEST)
sim = ns.Simulation(s, ctx)
for _ in range(3):
    sim.step()
f, c, n = sim.f, ctx, (s.nx, s.ny, s.nz)
tot = {"pre": 0.0, "pt": 0.0, "post": 0.0, "step": 0.0}
its = []
for rep in range(3):
    c.sync(); t0 = time.perf_counter()
    c.call("ns3d_update_tau", f["txx"], f["tyy"], f["tzz"], f["txy"], f["txz"], f["tyz"], f["Vx"], f["Vy"], f["Vz"], s.mu, s.dx, s.dy, s.dz, *n)
    c.call("ns3d_predict_V", f["Vx"], f["Vy"], f["Vz"], f["txx"], f["tyy"], f["tzz"], f["txy"], f["txz"], f["tyz"], s.rho, s.g, s.dt, s.dx, s.dy, s.dz, *n)
    sim.set_cylinder()
    c.call("ns3d_update_divV", f["divV"], f["Vx"], f["Vy"], f["Vz"], s.dx, s.dy, s.dz, *n)
    c.sync(); t1 = time.perf_counter()
    it, hist = c.pt_solve(f["Pr"], f["dPrdtau"], f["divV"], s.pt_params())
    c.sync(); t2 = time.perf_counter()
    c.call("ns3d_correct_V", f["Vx"], f["Vy"], f["Vz"], f["Pr"], s.dt, s.rho, s.dx, s.dy, s.dz, *n)
    sim.set_cylinder(); sim.set_bc_Vel()
    for a in ("Vx", "Vy", "Vz", "C"):
        c.copy(f[a + "_o"], f[a])
    c.call("ns3d_advect", f["Vx"], f["Vx_o"], f["Vy"], f["Vy_o"], f["Vz"], f["Vz_o"], f["C"], f["C_o"], s.dt, s.dx, s.dy, s.dz, *n)
    c.sync(); t3 = time.perf_counter()
    tot["pre"] += t1 - t0; tot["pt"] += t2 - t1; tot["post"] += t3 - t2; its.append((it, len(hist)))
print({k: round(v / 3 * 1e3, 3) for k, v in tot.items()}, its)
# pure iteration rate for the same count
import torch
stream = torch.cuda.ExternalStream(ctx.stream)
it = its[-1][0]
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record(stream); ctx.pt_iterate(f["Pr"], f["dPrdtau"], f["divV"], s.pt_params(), it); e1.record(stream); ctx.sync()
print("pt_iterate(%d) ms:" % it, e0.elapsed_time(e1), "-> us/iter", e0.elapsed_time(e1) / it * 1e3)
