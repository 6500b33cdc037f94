"""Burst vs sustained rate of the fused PT iteration, with 20 ms clock/power samples."""
import os, sys, subprocess, threading, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import navierstokes3d_b200 as ns
mode = sys.argv[1] if len(sys.argv) > 1 else "FASTEST"
g = sys.argv[2] if len(sys.argv) > 2 else "255x153x153"
nx, ny, nz = map(int, g.split("x"))
s = ns.setup_gpu(nx, ny=ny, nz=nz)
ctx = ns.Context(0, getattr(ns, mode))
stream = torch.cuda.ExternalStream(ctx.stream)
rng = np.random.default_rng(0)
Pr = ctx.from_host(np.asfortranarray(rng.uniform(-1, 1, size=(nx, ny, nz))))
dP = ctx.zeros(nx - 2, ny - 2, nz - 2)
dv = ctx.from_host(np.asfortranarray(rng.uniform(-1e-3, 1e-3, size=(nx, ny, nz))))
pt = s.pt_params(0)
lines = []
proc = subprocess.Popen(["nvidia-smi", "--query-gpu=clocks.sm,clocks.mem,power.draw,temperature.gpu", "--format=csv,noheader,nounits", "-lms", "20", "-i", "0"], stdout=subprocess.PIPE, text=True)
threading.Thread(target=lambda: [lines.append((time.perf_counter(), l.strip())) for l in proc.stdout], daemon=True).start()
n = 152 if nx < 400 else 16
ctx.pt_iterate(Pr, dP, dv, pt, n); ctx.sync()
time.sleep(1.0)
res = []
t_start = time.perf_counter()
for rep in range(int(sys.argv[3]) if len(sys.argv) > 3 else 60):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream); ctx.pt_iterate(Pr, dP, dv, pt, n); e1.record(stream); ctx.sync()
    res.append((time.perf_counter() - t_start, e0.elapsed_time(e1) / n * 1e3))
proc.terminate()
print(mode, g, "us/iter per chunk of", n, ":", [round(r[1], 1) for r in res[:3]], "...", [round(r[1], 1) for r in res[-3:]])
busy = [l for t, l in lines if t >= t_start]
print("samples during load (sm MHz, mem MHz, W, C):", busy[:2], "...", busy[len(busy)//2:len(busy)//2+2], "...", busy[-2:])
