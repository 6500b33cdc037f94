"""Where does the end-to-end time go?  One time step of config B alone, with an upload / download in flight on the
library's copy streams, and the copies alone (diagnostic for bench.py's e2e leg)."""
import os, sys, time, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import navierstokes3d_b200 as ns
from navierstokes3d_b200 import native as N

s = ns.setup_multi_gpu(255)
ctx = ns.Context(0, ns.FAST)
ns.driver.attach_communicator(ctx, 0, 1)
sim = ns.Simulation(s, ctx)
other = ns.Simulation(s, ctx)
for _ in range(3):
    sim.step()
snap = {k: sim.host(k) for k in ("Pr", "dPrdtau", "Vx", "Vy", "Vz", "C")}
pin = {k: torch.from_numpy(np.ascontiguousarray(v.ravel(order="F"))).pin_memory() for k, v in snap.items()}
out = {k: torch.empty_like(t).pin_memory() for k, t in pin.items()}
nbytes = sum(t.numel() * 8 for t in pin.values())

def reset():
    for k, v in snap.items():
        sim.f[k].set(v)
    ctx.sync()

def timed(fn, reps=3):
    best = 1e9
    for _ in range(reps):
        reset()
        t0 = time.perf_counter(); fn(); ctx.sync(); best = min(best, time.perf_counter() - t0)
    return best * 1e3

res = {"bytes_each_way": nbytes}
res["step_alone_ms"] = timed(lambda: sim.step())
def up_only():
    for k, t in pin.items(): ctx.h2d_async(other.f[k].ptr, t.data_ptr(), t.numel())
res["h2d_alone_ms"] = timed(up_only)
def down_only():
    for k, t in out.items(): ctx.d2h_async(t.data_ptr(), other.f[k].ptr, t.numel())
res["d2h_alone_ms"] = timed(down_only)
def step_with_up():
    up_only(); sim.step()
res["step_with_h2d_in_flight_ms"] = timed(step_with_up)
def step_with_both():
    up_only(); down_only(); sim.step()
res["step_with_h2d_and_d2h_in_flight_ms"] = timed(step_with_both)
def blocking():
    for k, t in pin.items(): ctx.h2d_raw(sim.f[k].ptr, t.data_ptr(), t.numel())
    sim.step()
    for k, t in out.items(): ctx.d2h_raw(t.data_ptr(), sim.f[k].ptr, t.numel())
res["blocking_ms"] = timed(blocking)
print(json.dumps(res))
ctx.close()
