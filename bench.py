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
import re
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
TEXT_RECORD = os.path.join(ROOT, "tests", "golden", "jl_reference_config_B.json")   # see the warm-up loop of main()
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
TOL = 1e-10   # north_star: "fields within a stated FP64 relative tolerance (e.g. 1e-10 after N steps)"; N = --steps


def kernel_key(workload: str, mode: str, opts) -> str:
    """Key of profiles/traffic.json: the ncu DRAM-traffic figure is only valid for the kernel source it
    was captured on, so the key carries a hash of the device code and every tuning option."""
    import hashlib
    h = hashlib.sha1()
    for f in ("ns3d_ptv_kernels.cuh", "ns3d_ptv.cu"):
        with open(os.path.join(ROOT, "navierstokes3d_b200", "csrc", f), "rb") as fh:
            h.update(fh.read())
    return f"{workload}:{mode}:{','.join(sorted(opts))}:{h.hexdigest()[:12]}"


def bind_numa(local: int) -> str:
    """Pin this rank to the CPUs next to its GPU (NVML affinity) BEFORE any pinned host buffer exists, so
    that first-touch places the e2e staging buffers on the GPU's NUMA node (round 1: all 8 ranks pinned on
    node 0 -> host copies 8.5 ms/step at N=1 but 35 ms/step at N=8)."""
    try:
        import pynvml
        pynvml.nvmlInit()
        h = pynvml.nvmlDeviceGetHandleByIndex(local)
        words = pynvml.nvmlDeviceGetCpuAffinity(h, (os.cpu_count() + 63) // 64)
        cpus = {64 * w + b for w, word in enumerate(words) for b in range(64) if (word >> b) & 1}
        cpus &= os.sched_getaffinity(0) if hasattr(os, "sched_getaffinity") else cpus
        if cpus:
            os.sched_setaffinity(0, cpus)
            return f"{len(cpus)} cpus near gpu{local}"
    except Exception as exc:  # noqa: BLE001
        return f"not bound ({type(exc).__name__})"
    return "not bound"


class Rig:
    """rank / world / barrier / reductions shared by every leg of the GPU arm."""

    def __init__(self, args):
        import torch
        self.torch = torch
        self.rank = int(os.environ.get("RANK", "0"))
        self.world = int(os.environ.get("WORLD_SIZE", "1"))
        self.local = int(os.environ.get("LOCAL_RANK", "0"))
        if self.world != args.gpus and self.world > 1:
            raise SystemExit(f"--gpus {args.gpus} but WORLD_SIZE={self.world}")
        self.numa = bind_numa(self.local)
        torch.cuda.set_device(self.local)
        self.dist = None
        if self.world > 1:
            import torch.distributed as dist
            dist.init_process_group("nccl", device_id=torch.device("cuda", self.local))
            self.dist = dist

    def barrier(self):
        if self.dist:
            self.dist.barrier()
        self.torch.cuda.synchronize()

    def reduce(self, vals, op: str):
        """element-wise max / min / sum over ranks of a list of floats"""
        if not self.dist:
            return [float(v) for v in vals]
        t = self.torch.tensor(list(vals), device="cuda", dtype=self.torch.float64)
        self.dist.all_reduce(t, op={"max": self.dist.ReduceOp.MAX, "min": self.dist.ReduceOp.MIN,
                                    "sum": self.dist.ReduceOp.SUM}[op])
        return [float(v) for v in t]

    def event(self):
        return self.torch.cuda.Event(enable_timing=True)


def make_sim(ns, rig, args, workload, variant, mode, fixed_iters=0):
    from navierstokes3d_b200.driver import attach_communicator
    _, nx, ny, nz = WORKLOADS[workload]
    kw = {}
    if fixed_iters:
        kw = dict(eps_it=0.0, niter=fixed_iters, nchk=max(fixed_iters // 2, 1))
    if variant == "G":
        s = ns.setup_gpu(nx, ny=ny, nz=nz, **kw)
    else:
        # weak scaling: the domain grows in z with the number of slabs so
        # that dz stays equal to dx and every GPU does the same work; the cylinder is z-invariant.
        probe = ns.setup_multi_gpu(nx, ny=ny, nz=nz, rank=rig.rank, nranks=rig.world, **kw)
        if rig.world > 1 or (ny is not None):
            kw = dict(kw, ly=probe.ny * probe.dx, lz=probe.grid.nz_g * probe.dx)
        s = ns.setup_multi_gpu(nx, ny=ny, nz=nz, rank=rig.rank, nranks=rig.world, **kw)
    ctx = ns.Context(rig.local, getattr(ns, mode))
    attach_communicator(ctx, rig.rank, rig.world)
    for o in args.opt:
        name, val = o.split("=")
        ctx.set_option(name, int(val))
    sim = ns.Simulation(s, ctx, zchunk=args.zchunk)
    stream = rig.torch.cuda.ExternalStream(ctx.stream, device=rig.local)
    return s, ctx, sim, stream


def timed_steps(rig, sim, stream, steps):
    """exactly `steps` time steps between CUDA events on the library's stream, barrier + sync on both sides;
    returns (seconds = max over ranks of max(device, wall), bytes summed over ranks, iters, checks) of THIS rank's run"""
    s = sim.s
    n_cells = s.nx * s.ny * s.nz
    rig.barrier()
    ev0, ev1 = rig.event(), rig.event()
    ev0.record(stream)
    t0 = time.perf_counter()
    iters, checks = [], []
    for _ in range(steps):
        it, hist = sim.step()
        iters.append(it)
        checks.append(len(hist))
    ev1.record(stream)
    rig.barrier()
    wall = time.perf_counter() - t0
    dev_s = ev0.elapsed_time(ev1) / 1e3
    bytes_rank = sum(a_eff_bytes(n_cells, i, c) for i, c in zip(iters, checks))
    t_all = rig.reduce([max(dev_s, wall)], "max")[0]
    bytes_all = rig.reduce([bytes_rank], "sum")[0]
    return t_all, bytes_all, iters, checks


STATE = ("Pr", "dPrdtau", "Vx", "Vy", "Vz", "C")


def field_diffs(a: dict, ref: dict) -> dict:
    """max |a - ref| per field over the field's scale; velocities share ONE scale (SURVEY.md "Hard parts":
    Vz is rounding noise in variant M, so a per-component relative error is meaningless)."""
    vscale = max(float(np.abs(ref[v]).max()) for v in ("Vx", "Vy", "Vz"))
    return {k: float(np.abs(a[k] - ref[k]).max() / (vscale if k[0] == "V" else max(float(np.abs(ref[k]).max()), 1e-300)))
            for k in a}


def extra_workload(ns, rig, args, workload, variant, mode, steps, warmup, fixed_iters):
    """one more configuration measured in the same run (device-resident T_eff only)"""
    s, ctx, sim, stream = make_sim(ns, rig, args, workload, variant, mode, fixed_iters)
    for _ in range(warmup):
        sim.step()
    t_all, bytes_all, iters, checks = timed_steps(rig, sim, stream, steps)
    out = {"workload": workload_name(workload, variant, s), "decomposition": f"z-slabs x{rig.world}", "mode": mode,
           "value": bytes_all / t_all / 1e9, "unit": "GB/s", "steps": steps, "warmup": warmup,
           "ms_per_step": t_all / steps * 1e3, "time_steps_per_s": steps / t_all,
           "t_eff_per_gpu": bytes_all / t_all / 1e9 / rig.world,
           "us_per_pt_iteration_whole_step": t_all / max(sum(iters), 1) * 1e6,
           "pt_iters_per_step": iters, "residual_checks_per_step": checks}
    ctx.close()
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="native", choices=["native", "reference"])
    ap.add_argument("--workload", default="B", choices=sorted(WORKLOADS))
    ap.add_argument("--mode", default="FAST", choices=["PARITY", "FAST", "FASTEST"],
                    help="arithmetic of the fused PT loop: FAST (default) is bit-identical to PARITY = the CPU oracle; "
                         "FASTEST trades that for FMA chains under the 1e-10 tolerance contract")
    ap.add_argument("--no-parity-check", action="store_true")
    ap.add_argument("--opt", action="append", default=[], help="library tuning option name=value (ns3d_set_option)")
    ap.add_argument("--zchunk", type=int, default=0)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-extras", action="store_true",
                    help="skip the extra legs of the default run (other arithmetic modes, variant M at N=1, config E)")
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

    import navierstokes3d_b200 as ns

    rig = Rig(args)
    rank, world, local = rig.rank, rig.world, rig.local
    torch = rig.torch
    barrier = rig.barrier

    variant = WORKLOADS[args.workload][0]
    if world > 1:
        variant = "M"   # the G script is single-GPU; z-slabs follow the multi-GPU script
    s, ctx, sim, stream = make_sim(ns, rig, args, args.workload, variant, args.mode, args.fixed_iters)
    _, nx, ny, nz = WORKLOADS[args.workload]
    n_cells = s.nx * s.ny * s.nz

    if args.pt_only:
        return pt_only(args, ns, ctx, s, stream, rank, world, barrier)

    # ---- device-resident timing: W warm-up steps, then exactly K steps -------------------------
    # The warm-up steps double as a check against the REFERENCE'S OWN TEXT: tests/golden/jl_reference_config_B.json holds
    # what `runme()` of scripts/NavierStokes3D_gpu.jl as shipped yields for its first time steps when oracle/jl_interp.py
    # executes the script line by line (iteration counts, every residual, SHA-256 of the fields).  Untimed.
    text_rec, text_check = None, None
    if world == 1 and args.workload == "B" and variant == "G" and not args.fixed_iters and not args.no_parity_check:
        try:
            with open(TEXT_RECORD) as fh:
                text_rec = json.load(fh)
            if text_rec.get("grid") != [s.nx, s.ny, s.nz]:
                text_rec = None
        except OSError:
            text_rec = None
    for w in range(args.warmup):
        it_w, hist_w = sim.step()
        if text_rec is not None and w < len(text_rec["steps"]):
            ref_w = text_rec["steps"][w]
            if text_check is None:
                text_check = {"against": "scripts/NavierStokes3D_gpu.jl as shipped, executed from its text (tests/golden/make_jl_config_B.py)",
                              "steps_checked": 0, "pt_iters_identical": True, "residuals_identical": True, "fields_bit_identical": None}
            text_check["steps_checked"] = w + 1
            text_check["pt_iters_identical"] &= (it_w == ref_w["iters"])
            if args.mode != "FASTEST":
                text_check["residuals_identical"] &= (list(hist_w) == ref_w["errs"])
                if w + 1 == min(args.warmup, len(text_rec["steps"])):
                    import hashlib
                    text_check["fields_bit_identical"] = all(
                        hashlib.sha256(np.asfortranarray(sim.host(k)).tobytes(order="F")).hexdigest() == ref_w["digest"][k]
                        for k in ("Pr", "Vx", "Vy", "Vz", "C"))
    snapshot = {k: sim.host(k) for k in STATE}   # for the e2e and parity legs: the same K steps again
    sampler = ClockSampler(local)
    sampler.start()
    l0 = ctx.launch_count
    t_all, bytes_all, iters, checks = timed_steps(rig, sim, stream, args.steps)
    clocks = sampler.stop()
    launches = ctx.launch_count - l0
    final = {k: sim.host(k) for k in ("Pr", "Vx", "Vy", "Vz", "C")} if not args.no_parity_check else None

    # ---- the dominant kernel, live: nchk fused PT iterations between events ----------------------
    pt = s.pt_params(args.zchunk)
    per_launch = ctx.pt_iters_per_launch(pt)
    n_probe = max(s.nchk, 48)
    n_probe -= n_probe % (2 * per_launch)   # the ping-pong buffers end where they started (graph cache hit)
    ctx.pt_iterate(sim.f["Pr"], sim.f["dPrdtau"], sim.f["divV"], pt, n_probe)   # untimed: builds this chunk's CUDA graph
    ctx.sync()
    barrier()
    k0, k1 = rig.event(), rig.event()
    k0.record(stream)
    ctx.pt_iterate(sim.f["Pr"], sim.f["dPrdtau"], sim.f["divV"], pt, n_probe)
    k1.record(stream)
    ctx.sync()
    desc = ctx.pt_kernel_name(pt)
    # ptv_flow_kernel (single rank): the n_probe iterations between the events are ONE launch of n_probe / per_launch
    # passes over the fields; ptv_kernel (z-slabs): one launch per pass of per_launch iterations
    persistent = desc.startswith("ptv_flow_kernel")
    iters_per_launch = n_probe if persistent else per_launch
    # z-bands: a pass over the fields is several launches on as many streams that overlap with the next pass's; what the
    # events measure -- and what `us_per_launch` then holds -- is the time per PASS
    m_bands = re.search(r"a pass = (\d+) launches", desc)
    launches_per_pass = int(m_bands.group(1)) if m_bands else 1
    t_launch = k0.elapsed_time(k1) / 1e3 / n_probe * iters_per_launch
    peak, peak_src = hbm_peak()
    achieved = 40.0 * iters_per_launch * n_cells / t_launch / 1e9
    traffic, tkey = None, kernel_key(args.workload, args.mode, args.opt)
    try:
        with open(os.path.join(ROOT, "profiles", "traffic.json")) as fh:
            traffic = json.load(fh).get(tkey)
    except Exception:  # noqa: BLE001
        pass

    # ---- end to end through the public API with HOST buffers ------------------------------------
    # Every step takes its input state from pinned host memory and returns its result there.  `e2e` runs that through
    # driver.StreamedSteps (two device field sets, uploads / downloads on the library's copy streams overlapped with
    # the neighbouring steps); `blocking_copies` is the same with ns3d_h2d / ns3d_d2h around every step.
    e2e = None
    if not args.no_e2e:
        pinned = {k: torch.from_numpy(np.ascontiguousarray(v.ravel(order="F"))).pin_memory() for k, v in snapshot.items()}
        pinned_out = {k: torch.empty_like(t).pin_memory() for k, t in pinned.items()}
        h2d_b = d2h_b = sum(t.numel() * 8 for t in pinned.values())
        # blocking copies (round 1's end-to-end figure)
        barrier()
        tw = time.perf_counter()
        b_iters, b_checks = [], []
        for _ in range(args.steps):
            for k, t in pinned.items():     # host -> device: this step's input state
                ctx.h2d_raw(sim.f[k].ptr, t.data_ptr(), t.numel())
            it, hist = sim.step()
            for k, t in pinned_out.items():     # device -> host: the step's result
                ctx.d2h_raw(t.data_ptr(), sim.f[k].ptr, t.numel())
            b_iters.append(it)
            b_checks.append(len(hist))
        barrier()
        b_t = rig.reduce([time.perf_counter() - tw], "max")[0]
        b_bytes = rig.reduce([sum(a_eff_bytes(n_cells, i, c) for i, c in zip(b_iters, b_checks))], "sum")[0]
        # pipelined
        runner = ns.StreamedSteps(s, ctx, zchunk=args.zchunk)
        ins = {k: (t.data_ptr(), t.numel()) for k, t in pinned.items()}
        outs = {k: (t.data_ptr(), t.numel()) for k, t in pinned_out.items()}
        runner.run(1, ins, outs)            # untimed: first touch of the second field set
        barrier()
        e0, e1 = rig.event(), rig.event()
        e0.record(stream)
        tw = time.perf_counter()
        res = runner.run(args.steps, ins, outs)
        e1.record(stream)
        barrier()
        e_wall = time.perf_counter() - tw
        e_dev = max(e0.elapsed_time(e1) / 1e3, e_wall)   # the copies end on other streams: wall clock is the honest one
        e_iters, e_checks = [r[0] for r in res], [len(r[1]) for r in res]
        e_bytes = sum(a_eff_bytes(n_cells, i, c) for i, c in zip(e_iters, e_checks))
        e_t = rig.reduce([e_dev], "max")[0]
        e_bytes_all = rig.reduce([e_bytes], "sum")[0]
        same_result = all(torch.equal(pinned_out[k], torch.from_numpy(np.ascontiguousarray(sim.host(k).ravel(order="F"))))
                          for k in ("Pr", "Vx", "C"))   # the pipelined run produced what the blocking run left on the device
        e2e = {"value": e_bytes_all / e_t / 1e9, "unit": "GB/s", "h2d_bytes_per_step": h2d_b, "d2h_bytes_per_step": d2h_b,
               "ms_per_step": e_t / args.steps * 1e3, "time_steps_per_s": args.steps / e_t, "host_numa": rig.numa,
               "how": "driver.StreamedSteps: pinned host state in and out every step, copies overlapped with the neighbouring steps",
               "pt_iters_per_step": e_iters, "matches_blocking_run": bool(same_result),
               "blocking_copies": {"value": b_bytes / b_t / 1e9, "unit": "GB/s", "ms_per_step": b_t / args.steps * 1e3}}
        runner.close()

    # ---- parity of the timed steps: the same K steps again in PARITY mode (bit-equal to the CPU oracle,
    # tests/test_gpu_solver.py): PT iteration counts must be identical on EVERY rank and every field of EVERY
    # rank's slab within TOL of the common scale after these K steps.  A miss makes the run fail (rc 3). ----
    parity, parity_ok = None, True
    if final is not None:
        def rerun(mode):
            for k, v in snapshot.items():
                sim.f[k].set(v)
            ctx.set_mode(getattr(ns, mode))
            ctx.sync()
            barrier()
            t0 = time.perf_counter()
            res = [sim.step() for _ in range(args.steps)]
            ctx.sync()
            barrier()
            dt = time.perf_counter() - t0
            ctx.set_mode(getattr(ns, args.mode))
            return res, dt, {k: sim.host(k) for k in final}

        p_res, _, ref = rerun("PARITY")
        p_iters = [r[0] for r in p_res]
        diffs = field_diffs(final, ref)
        names = sorted(diffs)
        worst = dict(zip(names, rig.reduce([diffs[k] for k in names], "max")))
        same = rig.reduce([float(p_iters == iters)], "min")[0] == 1.0
        within = max(worst.values()) <= TOL
        parity_ok = same and within
        parity = {"against": "the same steps in PARITY mode (IEEE division, no FMA; bit-equal to the CPU oracle in tests/)",
                  "horizon_steps": args.steps, "ranks_checked": world,
                  "pt_iters_identical": same, "max_rel_diff": worst, "tolerance": TOL, "within_tolerance": within,
                  "bit_identical": max(worst.values()) == 0.0}
        if not args.no_extras:
            # what the other arithmetic costs / buys: the same steps once more in the other non-PARITY mode
            other = "FASTEST" if args.mode != "FASTEST" else "FAST"
            o_res, o_dt, o_fields = rerun(other)
            o_dt = rig.reduce([o_dt], "max")[0]
            o_bytes = rig.reduce([sum(a_eff_bytes(n_cells, it, len(h)) for it, h in o_res)], "sum")[0]
            od = field_diffs(o_fields, ref)
            o_worst = dict(zip(names, rig.reduce([od[k] for k in names], "max")))
            parity["other_mode"] = {"mode": other, "value": o_bytes / o_dt / 1e9, "unit": "GB/s (wall clock, max over ranks)",
                                    "pt_iters_identical": rig.reduce([float([r[0] for r in o_res] == p_iters)], "min")[0] == 1.0,
                                    "max_rel_diff_vs_parity": o_worst, "within_tolerance": max(o_worst.values()) <= TOL}

    share = (sum(iters) / args.steps) * (t_launch / iters_per_launch) / (t_all / args.steps)
    ctx.close()

    # ---- extra legs of the default run (each on a fresh context; the B fields are freed above) ----
    extras = {}
    if not args.no_extras and args.workload == "B" and not args.fixed_iters:
        if world == 1:
            # like-with-like base of the weak-scaling curve: N > 1 runs variant M (the multi-GPU script),
            # so here is variant M on ONE GPU, same grid, same step counts
            extras["variant_M_n1"] = extra_workload(ns, rig, args, "B", "M", args.mode, args.steps, args.warmup, 0)
        # BASELINE configs[4]: 511^3 cells per GPU, fixed PT work (eps_it = 0, 1020 iterations,
        # residual check + allreduce every 510), nt = 3
        extras["config_E"] = extra_workload(ns, rig, args, "E", "M", args.mode, 3, 1, 1020)

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
            "roofline": {"bound": "hbm", "kernel": desc,
                         "achieved": achieved,
                         "peak": peak, "peak_source": peak_src, "unit": "GB/s", "frac": achieved / peak,
                         "traffic": traffic, "traffic_key": tkey, "us_per_launch": t_launch * 1e6,
                         "pt_iterations_per_launch": iters_per_launch, "pt_iterations_per_pass": per_launch,
                         "launches_per_pass": launches_per_pass,
                         "us_per_pass": t_launch * 1e6 / iters_per_launch * per_launch,
                         # what actually crossed the DRAM interface (ncu) over the same launch time: the
                         # kernel keeps the intermediate iterates on chip, so this is well below `achieved`
                         "dram_achieved": (traffic / t_launch / 1e9) if traffic else None,
                         "dram_frac": (traffic / t_launch / 1e9 / peak) if traffic else None,
                         "algorithmic_bytes_per_launch": 40.0 * iters_per_launch * n_cells,
                         "share_of_step": share},
        }
        if parity:
            line["parity_check"] = parity
            if text_check is not None:
                parity["reference_text"] = text_check
        if e2e:
            line["e2e"] = e2e
        line.update(extras)
        if world == 1 and not args.no_cpu_baseline:
            cv, cores, sample, _, _ = oracle_sample(variant, nx, ny, nz)
            line["cpu_baseline"] = {"value": cv, "unit": "GB/s", "cores": cores, "kind": "port", "sample": sample}
        print(json.dumps(line))
    if world > 1:
        rig.dist.destroy_process_group()
    if text_check is not None and not (text_check["pt_iters_identical"] and text_check["residuals_identical"]
                                       and text_check["fields_bit_identical"] in (True, None)):
        parity_ok = False
    if not parity_ok:
        sys.stderr.write("bench.py: PARITY CHECK FAILED (iteration counts differ or a field is outside the tolerance)\n")
        sys.exit(3)


if __name__ == "__main__":
    main()
