#!/usr/bin/env python
"""bench.py -- headline benchmark of the NavierStokes3D hot path on B200.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference] [--workload B|C|E] [--mode ...]

Metric (BASELINE.json): T_eff [GB/s] (ParallelStencil convention) and time-steps/s of the
cylinder flow at 255x153x153 cells per GPU; weak scaling over z-slabs for N > 1.

A "step" is one time step of the solver (M:449-477 / G:121-142): predictor, the whole
pseudo-transient pressure loop with its residual checks, corrector, advection.
Algorithmic bytes per step:  A_eff = (21 + 5*N_iter + 2*N_chk) * 8 * nx*ny*nz.

Printed keys beyond the driver contract: "time_steps_per_s", "pt_iters_per_step",
"roofline" (fused PT kernel: 40 B/cell/iteration, two iterations per launch, over the measured HBM peak),
"cpu_baseline" (the CPU oracle, a restated reference CPU path, timed on a bounded sample).
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

WORKLOADS = {
    # name: (variant at N=1, local nx, ny, nz)  -- ny/nz None = the script's ceil(0.6*nx) rule
    "B": ("G", 255, None, None),     # BASELINE configs[1]: scripts/NavierStokes3D_gpu.jl as shipped
    "C": ("M", 511, 511, 511),       # configs[2] geometry (Poisson-only timing uses --pt-only)
    "E": ("M", 511, 511, 511),       # configs[4]: 511^3 per GPU
    "A": ("M", 63, None, None),      # configs[0] (parity grid; launch-bound, for completeness)
    "D": ("M", 1023, 511, 511),      # configs[3]: large single-GPU cylinder flow (2.14 GB per field)
}


def a_eff_bytes(n_cells: int, n_iter: int, n_chk: int) -> float:
    return (21 + 5 * n_iter + 2 * n_chk) * 8.0 * n_cells


def workload_name(key: str, variant: str, s) -> str:
    """config.workload, shared by both arms (s: anything with nx, ny, nz, eps_it, niter, nchk)."""
    script = "scripts/NavierStokes3D_gpu.jl" if variant == "G" else "scripts/NavierStokes3D_multi_gpu.jl"
    return (f"{key}: cylinder flow {s.nx}x{s.ny}x{s.nz} cells per GPU, Float64, variant {variant} ({script} parameters), "
            f"eps_it={s.eps_it}, niter={s.niter}, nchk={s.nchk}")


def hbm_peak():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    try:
        with open(path) as fh:
            return float(json.load(fh)["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    except Exception:  # noqa: BLE001
        return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler:
    """nvidia-smi clocks + throttle reasons during the timed region (B200_PROFILING.md recipe)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
         "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index: int):
        self.index, self.proc, self.lines = index, None, []

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-lms", "200", "-i", str(self.index)], stdout=subprocess.PIPE, text=True)
            threading.Thread(target=self._pump, daemon=True).start()
        except Exception:  # noqa: BLE001
            self.proc = None

    def _pump(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self) -> dict:
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.25)
        self.proc.terminate()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ln in self.lines:
            parts = [x.strip() for x in ln.split(",")]
            if len(parts) < 9:
                continue
            try:
                sm.append(float(parts[1]))
                mx.append(float(parts[2]))
            except ValueError:
                continue
            for nm, val in zip(names, parts[5:9]):
                if val.lower().startswith("active"):
                    reasons.add(nm)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


# ------------------------------------------------------------------------------------------------
# CPU oracle legs (the only places bench.py touches oracle/)
# ------------------------------------------------------------------------------------------------
def oracle_sample(variant: str, nx: int, ny, nz, budget_s: float = 12.0):
    """Times the restated reference CPU path (oracle/ns3d_oracle.c, OpenMP, all host cores) on a
    bounded sample of the workload: one time step whose PT loop is cut after `n_it` iterations
    (one residual check included), n_it chosen from a probe so the sample is ~budget_s."""
    from oracle import oracle as O
    p = O.params_G(nx, ny=ny, nz=nz) if variant == "G" else O.params_M(nx, ny=ny, nz=nz)
    # all the host threads the box offers (torchrun exports OMP_NUM_THREADS=1 to its ranks)
    O.lib().ns3d_oracle_set_num_threads(len(os.sched_getaffinity(0)))
    cores = O.lib().ns3d_oracle_num_threads()
    f = O.initial_fields(p)
    t0 = time.perf_counter()
    for _ in range(2):
        O.update_dPrdtau(p, f); O.update_Pr(p, f); O.set_bc_Pr(p, f)
    t_it = (time.perf_counter() - t0) / 2
    n_it = int(max(4, min(20 * p.nchk, budget_s / max(t_it, 1e-6))))
    f = O.initial_fields(p)
    p.niter, p.nchk = n_it, n_it
    t0 = time.perf_counter()
    iters, hist = O.step(p, f)
    dt = time.perf_counter() - t0
    n = p.nx * p.ny * p.nz
    teff = a_eff_bytes(n, iters, len(hist)) / dt / 1e9
    sample = (f"1 time step of {p.nx}x{p.ny}x{p.nz} variant {variant} with the PT loop cut at {iters} iterations "
              f"+ {len(hist)} residual check ({dt:.1f} s)")
    return teff, cores, sample, dt, iters


def run_reference(args):
    """--impl reference: the reference's own CPU implementation of the path.  Julia is absent
    (no ParallelStencil Threads backend can run), so this is the oracle port, all host threads."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    variant, nx, ny, nz = WORKLOADS[args.workload]
    vals, last = [], None
    budget = args.sample_seconds
    for i in range(args.warmup + args.steps):
        teff, cores, sample, dt, iters = oracle_sample(variant, nx, ny, nz, budget_s=budget)
        if i >= args.warmup:
            vals.append((teff, dt, iters))
        last = (cores, sample)
    teff = float(np.mean([v[0] for v in vals]))
    ms = float(np.mean([v[1] for v in vals]) * 1e3)
    from oracle import oracle as O
    p = O.params_G(nx, ny=ny, nz=nz) if variant == "G" else O.params_M(nx, ny=ny, nz=nz)
    line = {
        "impl": "reference", "metric": "T_eff", "value": teff, "unit": "GB/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": workload_name(args.workload, variant, p), "decomposition": "z-slabs x1",
                   "implementation": "CPU oracle port of the reference's kernels, OpenMP over z like ParallelStencil's "
                                     "Threads backend (Julia / ParallelStencil are not installable here); each step is "
                                     "a bounded sample: one time step with the PT loop cut short"},
        "cpu_baseline": {"value": teff, "unit": "GB/s", "cores": last[0], "kind": "port", "sample": last[1]},
        "e2e": {"value": teff, "unit": "GB/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line))


# ------------------------------------------------------------------------------------------------
# Poisson-only leg (BASELINE configs[2])
# ------------------------------------------------------------------------------------------------
def pt_only(args, ns, ctx, s, stream, rank, world, barrier):
    import torch
    nx, ny, nz = s.nx, s.ny, s.nz
    n = nx * ny * nz
    rng = np.random.default_rng(1234)
    x = (np.arange(nx) + 0.5) / nx
    y = (np.arange(ny) + 0.5) / ny
    z = (np.arange(nz) + 0.5) / nz
    rhs = (np.sin(2 * np.pi * x)[:, None, None] * np.sin(2 * np.pi * y)[None, :, None]
           * np.sin(2 * np.pi * z)[None, None, :])
    rhs += rng.uniform(-0.01, 0.01, size=rhs.shape)
    Pr = ctx.zeros(nx, ny, nz)
    dP = ctx.zeros(nx - 2, ny - 2, nz - 2)
    dv = ctx.from_host(np.asfortranarray(rhs))
    del rhs
    pt = s.pt_params(args.zchunk)
    pt.eps_it, pt.niter, pt.nchk = 0.0, args.pt_only, min(510, args.pt_only)
    ctx.pt_solve(Pr, dP, dv, pt)          # warm-up (also builds the graph)
    times = []
    for _ in range(max(args.steps, 1)):
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        iters, hist = ctx.pt_solve(Pr, dP, dv, pt)
        e1.record(stream)
        barrier()
        times.append(e0.elapsed_time(e1) / 1e3)
    t = float(np.median(times))
    peak, peak_src = hbm_peak()
    if rank == 0:
        bytes_ = (5 * iters + 2 * len(hist)) * 8.0 * n * world
        print(json.dumps({"metric": "T_eff", "value": bytes_ / t / 1e9, "unit": "GB/s", "n_gpus": world,
                          "steps": args.steps, "warmup": 1, "ms_per_step": t * 1e3, "higher_is_better": True,
                          "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
                          "config": {"workload": f"Poisson-only PT solve {nx}x{ny}x{nz} per GPU, {iters} iterations, "
                                                 f"{len(hist)} residual checks, rhs = sin*sin*sin + U(-0.01,0.01) (seed 1234)",
                                     "mode": args.mode},
                          "us_per_iteration": t / iters * 1e6, "frac_of_hbm_peak_per_gpu": bytes_ / t / 1e9 / world / peak,
                          "final_err": hist[-1] if hist else None}))
    ctx.close()


# ------------------------------------------------------------------------------------------------
# GPU arm
# ------------------------------------------------------------------------------------------------
def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="native", choices=["native", "reference"])
    ap.add_argument("--workload", default="B", choices=sorted(WORKLOADS))
    ap.add_argument("--mode", default="FASTEST", choices=["PARITY", "FAST", "FASTEST"])
    ap.add_argument("--no-parity-check", action="store_true")
    ap.add_argument("--opt", action="append", default=[], help="library tuning option name=value (ns3d_set_option)")
    ap.add_argument("--zchunk", type=int, default=0)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--sample-seconds", type=float, default=8.0,
                    help="--impl reference: CPU seconds of the bounded sample that makes one step")
    ap.add_argument("--pt-only", type=int, default=0,
                    help="Poisson-only benchmark (BASELINE configs[2]): time exactly this many PT iterations "
                         "(ns3d_pt_solve with eps_it=0, one residual check per 510) on a synthetic rhs instead of time steps")
    ap.add_argument("--fixed-iters", type=int, default=0,
                    help="fixed PT work per step (eps_it=0, niter=this, nchk=niter/2): SURVEY.md 8d config E")
    args = ap.parse_args()
    if args.impl == "reference":
        return run_reference(args)

    import torch

    import navierstokes3d_b200 as ns
    from navierstokes3d_b200.driver import attach_communicator

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world != args.gpus and world > 1:
        raise SystemExit(f"--gpus {args.gpus} but WORLD_SIZE={world}")
    torch.cuda.set_device(local)
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    variant, nx, ny, nz = WORKLOADS[args.workload]
    if world > 1:
        variant = "M"   # the G script is single-GPU; z-slabs follow the multi-GPU script
    kw = {}
    if args.fixed_iters:
        kw = dict(eps_it=0.0, niter=args.fixed_iters, nchk=max(args.fixed_iters // 2, 1))
    if variant == "G":
        s = ns.setup_gpu(nx, ny=ny, nz=nz, **kw)
    else:
        # weak scaling: the domain grows in z with the number of slabs so
        # that dz stays equal to dx and every GPU does the same work; the cylinder is z-invariant.
        probe = ns.setup_multi_gpu(nx, ny=ny, nz=nz, rank=rank, nranks=world, **kw)
        if world > 1 or (ny is not None):
            kw = dict(kw, ly=probe.ny * probe.dx, lz=probe.grid.nz_g * probe.dx)
        s = ns.setup_multi_gpu(nx, ny=ny, nz=nz, rank=rank, nranks=world, **kw)
    ctx = ns.Context(local, getattr(ns, args.mode))
    attach_communicator(ctx, rank, world)
    for o in args.opt:
        name, val = o.split("=")
        ctx.set_option(name, int(val))
    sim = ns.Simulation(s, ctx, zchunk=args.zchunk)
    n_cells = s.nx * s.ny * s.nz
    stream = torch.cuda.ExternalStream(ctx.stream, device=local)

    if args.pt_only:
        return pt_only(args, ns, ctx, s, stream, rank, world, barrier)

    # ---- device-resident timing: W warm-up steps, then exactly K steps -------------------------
    for _ in range(args.warmup):
        sim.step()
    state_names = ("Pr", "dPrdtau", "Vx", "Vy", "Vz", "C")
    snapshot = {k: sim.host(k) for k in state_names}   # for the e2e leg: same K steps again
    sampler = ClockSampler(local)
    barrier()
    sampler.start()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    l0 = ctx.launch_count
    ev0.record(stream)
    t0 = time.perf_counter()
    iters, checks = [], []
    for _ in range(args.steps):
        it, hist = sim.step()
        iters.append(it)
        checks.append(len(hist))
    ev1.record(stream)
    barrier()
    wall = time.perf_counter() - t0
    clocks = sampler.stop()
    launches = ctx.launch_count - l0
    dev_s = ev0.elapsed_time(ev1) / 1e3
    bytes_rank = sum(a_eff_bytes(n_cells, i, c) for i, c in zip(iters, checks))
    final = {k: sim.host(k) for k in ("Pr", "Vx", "Vy", "Vz", "C")} if not args.no_parity_check else None

    # ---- the dominant kernel, live: nchk fused PT iterations between events ----------------------
    n_probe = max(s.nchk, 50) & ~1   # even: the ping-pong buffers end where they started (graph cache hit)
    pt = s.pt_params(args.zchunk)
    ctx.pt_iterate(sim.f["Pr"], sim.f["dPrdtau"], sim.f["divV"], pt, n_probe)   # untimed: builds this chunk's CUDA graph
    ctx.sync()
    barrier()
    k0, k1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    k0.record(stream)
    ctx.pt_iterate(sim.f["Pr"], sim.f["dPrdtau"], sim.f["divV"], pt, n_probe)
    k1.record(stream)
    ctx.sync()
    t_launch = k0.elapsed_time(k1) / 1e3 / n_probe
    peak, peak_src = hbm_peak()
    tb2_on = "tb2=0" not in args.opt     # default path: pt_tb2_kernel, two PT iterations per launch
    per_launch = 2 if tb2_on else 1
    t_launch *= per_launch
    achieved = 40.0 * per_launch * n_cells / t_launch / 1e9
    traffic = None
    try:
        with open(os.path.join(ROOT, "profiles", "traffic.json")) as fh:
            traffic = json.load(fh).get(f"{args.workload}:{args.mode}" + (":tb2" if tb2_on else ""))
    except Exception:  # noqa: BLE001
        pass

    # ---- end to end through the public API with HOST buffers ------------------------------------
    e2e = None
    if not args.no_e2e:
        pinned = {k: torch.from_numpy(np.ascontiguousarray(v.ravel(order="F"))).pin_memory() for k, v in snapshot.items()}
        h2d_b = d2h_b = sum(t.numel() * 8 for t in pinned.values())
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        tw = time.perf_counter()
        e_iters, e_checks = [], []
        for _ in range(args.steps):
            for k, t in pinned.items():     # host -> device: this step's input state
                ctx.h2d_raw(sim.f[k].ptr, t.data_ptr(), t.numel())
            it, hist = sim.step()
            for k, t in pinned.items():     # device -> host: the step's result
                ctx.d2h_raw(t.data_ptr(), sim.f[k].ptr, t.numel())
            e_iters.append(it)
            e_checks.append(len(hist))
        e1.record(stream)
        barrier()
        e_wall = time.perf_counter() - tw
        e_dev = max(e0.elapsed_time(e1) / 1e3, e_wall)   # host-blocking copies: wall clock is the honest one
        e_bytes = sum(a_eff_bytes(n_cells, i, c) for i, c in zip(e_iters, e_checks))
        e2e = [e_bytes, e_dev, h2d_b, d2h_b, e_iters]

    # ---- parity of the timed steps: the same K steps again in PARITY mode (bit-equal to the CPU
    # oracle, tests/test_gpu_solver.py): iteration counts must be identical, fields within 1e-10 ----
    parity = None
    if final is not None:
        for k, v in snapshot.items():
            sim.f[k].set(v)
        ctx.set_mode(ns.PARITY)
        p_iters = [sim.step()[0] for _ in range(args.steps)]
        ctx.set_mode(getattr(ns, args.mode))
        ref = {k: sim.host(k) for k in final}
        vscale = max(float(np.abs(ref[v]).max()) for v in ("Vx", "Vy", "Vz"))
        diffs = {k: float(np.abs(final[k] - ref[k]).max() / (vscale if k[0] == "V" else max(float(np.abs(ref[k]).max()), 1e-300)))
                 for k in final}
        parity = {"against": "the same steps in PARITY mode (IEEE division, no FMA; bit-equal to the CPU oracle in tests/)",
                  "pt_iters_identical": p_iters == iters, "max_rel_diff": diffs, "tolerance": 1e-10,
                  "within_tolerance": max(diffs.values()) <= 1e-10}
        if args.mode == "FASTEST":
            # the same steps once more in FAST mode (reference arithmetic via corrected reciprocal
            # division): what the bit-identical path costs
            for k, v in snapshot.items():
                sim.f[k].set(v)
            ctx.set_mode(ns.FAST)
            ctx.sync()
            tf = time.perf_counter()
            f_res = [sim.step() for _ in range(args.steps)]
            ctx.sync()
            tf = time.perf_counter() - tf
            ctx.set_mode(ns.FASTEST)
            f_bytes = sum(a_eff_bytes(n_cells, it, len(h)) for it, h in f_res)
            fdiff = max(float(np.abs(sim.host(k) - ref[k]).max()) for k in final)
            parity["fast_mode"] = {"value": f_bytes / tf / 1e9 * world, "unit": "GB/s (per-rank wall clock x ranks)",
                                   "pt_iters_identical": [r[0] for r in f_res] == p_iters,
                                   "max_abs_diff_vs_parity": fdiff}

    # ---- reduce over ranks: max time, summed bytes ----------------------------------------------
    t_rank = max(dev_s, wall)
    if world > 1:
        v = torch.tensor([t_rank, e2e[1] if e2e else 0.0], device="cuda", dtype=torch.float64)
        dist.all_reduce(v, op=dist.ReduceOp.MAX)
        b = torch.tensor([bytes_rank, e2e[0] if e2e else 0.0], device="cuda", dtype=torch.float64)
        dist.all_reduce(b, op=dist.ReduceOp.SUM)
        t_all, e_t = float(v[0]), float(v[1])
        bytes_all, e_bytes_all = float(b[0]), float(b[1])
    else:
        t_all, bytes_all = t_rank, bytes_rank
        e_t, e_bytes_all = (e2e[1], e2e[0]) if e2e else (0.0, 0.0)

    if rank == 0:
        teff = bytes_all / t_all / 1e9
        line = {
            "metric": "T_eff", "value": teff, "unit": "GB/s", "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": t_all / args.steps * 1e3, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {
                "workload": workload_name(args.workload, variant, s),
                "decomposition": f"z-slabs x{world}", "mode": args.mode,
                "l2_policy": "inputs larger than L2: PT working set 4 fields x %.1f MB > 126 MB" % (n_cells * 8 / 1e6)
                if 4 * n_cells * 8 > 126e6 else "working set fits L2 (not a bandwidth figure)",
            },
            "time_steps_per_s": args.steps / t_all * 1.0,
            "pt_iters_per_step": iters, "residual_checks_per_step": checks,
            "t_eff_per_gpu": teff / world, "frac_of_hbm_peak_per_gpu": teff / world / peak,
            "gpu_launches": int(launches),
            "clocks": clocks,
            "roofline": {"bound": "hbm",
                         "kernel": ("pt_tb2s_kernel (two fused PT iterations per launch: 2 x (K5+K6+set_bc_Pr!)"
                                    + ("; slab-interface chunks: pt_tb2_kernel<.,16,true>)" if world > 1 else ")") if tb2_on
                                    else "pt_iter_kernel (fused K5+K6+set_bc_Pr!)"),
                         "achieved": achieved,
                         "peak": peak, "peak_source": peak_src, "unit": "GB/s", "frac": achieved / peak,
                         "traffic": traffic, "us_per_launch": t_launch * 1e6,
                         # what actually crossed the DRAM interface (ncu) over the same launch time: the
                         # kernel keeps Pr^(1) and dPrdtau^(1) on chip, so this is well below `achieved`
                         "dram_achieved": (traffic / t_launch / 1e9) if traffic else None,
                         "dram_frac": (traffic / t_launch / 1e9 / peak) if traffic else None,
                         "algorithmic_bytes_per_launch": 40.0 * per_launch * n_cells,
                         "share_of_step": (sum(iters) / args.steps) * (t_launch / per_launch) / (t_all / args.steps)},
        }
        if parity:
            line["parity_check"] = parity
        if e2e:
            line["e2e"] = {"value": e_bytes_all / e_t / 1e9, "unit": "GB/s", "h2d_bytes_per_step": e2e[2],
                           "d2h_bytes_per_step": e2e[3], "ms_per_step": e_t / args.steps * 1e3,
                           "time_steps_per_s": args.steps / e_t}
        if world == 1 and not args.no_cpu_baseline:
            cv, cores, sample, _, _ = oracle_sample(variant, nx, ny, nz)
            line["cpu_baseline"] = {"value": cv, "unit": "GB/s", "cores": cores, "kind": "port", "sample": sample}
        print(json.dumps(line))
    ctx.close()
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
