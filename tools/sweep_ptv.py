"""Times the fused PT loop (ptv_flow_kernel / ptv_kernel on the pitched copies) over iterations per launch, rows per thread,
launch bounds, tile shapes and chunk lengths (tuning aid; results go to profiles/*.jsonl).

    python tools/sweep_ptv.py --grids 255x153x153 --modes FAST,FASTEST [--quick]
"""
import argparse, json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import navierstokes3d_b200 as ns

ap = argparse.ArgumentParser()
ap.add_argument("--grids", default="255x153x153")
ap.add_argument("--modes", default="FAST,FASTEST")
ap.add_argument("--iters", type=int, default=120)
ap.add_argument("--reps", type=int, default=3)
ap.add_argument("--sets", default="", help="';'-separated sets of ','-separated name=value options; zchunk=N is the chunk length")
args = ap.parse_args()


def default_sets():
    out = []
    for k in (2, 3, 1):
        for lb in (0, 1, 2, 3, 4):
            out.append({"ptv_k": k, "ptv_lb": lb})
    return out


sets = [dict((kv.split("=")[0], int(kv.split("=")[1])) for kv in s.split(",") if kv) for s in args.sets.split(";")] if args.sets else default_sets()
rng = np.random.default_rng(0)
for g in args.grids.split(","):
    nx, ny, nz = map(int, g.split("x"))
    s = ns.setup_gpu(nx, ny=ny, nz=nz)
    n = nx * ny * nz
    host_pr = np.asfortranarray(rng.uniform(-1, 1, size=(nx, ny, nz)))
    host_dv = np.asfortranarray(rng.uniform(-1e-3, 1e-3, size=(nx, ny, nz)))
    for mode in args.modes.split(","):
        ctx = ns.Context(0, getattr(ns, mode))
        stream = torch.cuda.ExternalStream(ctx.stream)
        Pr = ctx.from_host(host_pr)
        dP = ctx.zeros(nx - 2, ny - 2, nz - 2)
        dv = ctx.from_host(host_dv)
        for opts in sets:
            for name in ("ptv_k", "ptv_ns", "ptv_pxt", "ptv_bty"):
                ctx.set_option(name, 0)
            ctx.set_option("ptv_lb", -1)
            ctx.set_option("ptv_tma", 1)
            ctx.set_option("ptv_flow", 0)
            ctx.set_option("ptv_bands", -1)
            ctx.set_option("serpentine", -1)
            zc = 0
            for k, v in opts.items():
                if k == "zchunk":
                    zc = v
                else:
                    ctx.set_option(k, v)
            pt = s.pt_params(zc)
            try:
                desc = ctx.pt_kernel_name(pt)
                per = ctx.pt_iters_per_launch(pt)
                iters = args.iters - args.iters % (2 * per)
                ctx.pt_iterate(Pr, dP, dv, pt, iters)   # warm-up: builds the CUDA graph of this chunk
                ctx.sync()
                best = 1e9
                for rep in range(args.reps):
                    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                    e0.record(stream)
                    ctx.pt_iterate(Pr, dP, dv, pt, iters)
                    e1.record(stream)
                    ctx.sync()
                    best = min(best, e0.elapsed_time(e1) / iters * 1e3)
                # pack + unpack of a pt_iterate call are inside the timing: 10 field passes per `iters` iterations
                print(json.dumps({"grid": g, "mode": mode, "opts": opts, "us_per_iter": round(best, 2),
                                  "T_eff_GBs": round(40.0 * n / best / 1e3, 1), "kernel": desc.split(" (")[0],
                                  "shape": desc.split("tiles of ")[-1].rstrip(")") if "tiles of" in desc else ""}), flush=True)
            except Exception as exc:  # noqa: BLE001
                print(json.dumps({"grid": g, "mode": mode, "opts": opts, "error": str(exc)[:200]}), flush=True)
        ctx.close()
