"""Times the fused PT iteration (us/iteration, T_eff) over modes / zchunk / grid sizes / library
options (tuning aid; results go to profiles/*.jsonl).

    python tools/sweep_pt.py --grids 255x153x153 --modes FASTEST --zchunks 0,19 --tb2ty 16 \
        --sets "tb2_slim=0;tb2_slim=1;tb2_slim=1,tb2_pf=1"
"""
import argparse, json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import navierstokes3d_b200 as ns

ap = argparse.ArgumentParser()
ap.add_argument("--grids", default="255x153x153,511x511x511")
ap.add_argument("--modes", default="PARITY,FAST,FASTEST")
ap.add_argument("--zchunks", default="0,2,4,8,16,32")
ap.add_argument("--iters", type=int, default=200)
ap.add_argument("--reps", type=int, default=3)
ap.add_argument("--variant", default="G")
ap.add_argument("--minb", default="0")
ap.add_argument("--serp", default="-1")
ap.add_argument("--tb2", default="1")
ap.add_argument("--tb2ty", default="16")
ap.add_argument("--sets", default="", help="';'-separated sets of ','-separated name=value library options")
args = ap.parse_args()
rng = np.random.default_rng(0)
sets = [dict(kv.split("=") for kv in s.split(",") if kv) for s in args.sets.split(";")] if args.sets else [{}]
for g in args.grids.split(","):
    nx, ny, nz = map(int, g.split("x"))
    s = ns.setup_gpu(nx, ny=ny, nz=nz) if args.variant == "G" else ns.setup_multi_gpu(nx, ny=ny, nz=nz)
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
            for tbty in map(int, args.tb2ty.split(",")):
                for sp in map(int, args.serp.split(",")):
                    for minb in map(int, args.minb.split(",")):
                        for zc in map(int, args.zchunks.split(",")):
                            ctx.set_option("tb2_ty", tbty)
                            ctx.set_option("pt_minb", minb)
                            ctx.set_option("serpentine", sp)
                            ctx.set_option("tb2", int(args.tb2))
                            for k, v in opts.items():
                                ctx.set_option(k, int(v))
                            pt = s.pt_params(zc)
                            ctx.pt_iterate(Pr, dP, dv, pt, args.iters)   # warm-up: builds the CUDA graph of this chunk
                            ctx.sync()
                            best = 1e9
                            for rep in range(args.reps):
                                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                                e0.record(stream)
                                ctx.pt_iterate(Pr, dP, dv, pt, args.iters)
                                e1.record(stream)
                                ctx.sync()
                                best = min(best, e0.elapsed_time(e1) / args.iters * 1e3)
                            print(json.dumps({"grid": g, "mode": mode, "zchunk": zc, "minb": minb, "serp": sp,
                                              "tb2": int(args.tb2), "tb2_ty": tbty, "opts": opts,
                                              "us_per_iter": round(best, 2),
                                              "T_eff_GBs": round(40.0 * n / best / 1e3, 1)}), flush=True)
        ctx.close()
