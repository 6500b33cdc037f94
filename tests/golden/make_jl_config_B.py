"""The benchmark configuration itself, from the reference's text: `runme()` of scripts/NavierStokes3D_gpu.jl
EXACTLY AS SHIPPED (nx = 255 -> 255x153x153, g = 9.81, eps = 1e-3, nchk = 152, G:15-61) executed by
oracle/jl_interp.py for the first NT time steps (only the literal `nt = 10000`, G:51, is replaced).

    python tests/golden/make_jl_config_B.py [NT] [G|M]   (needs /root/reference; about half an hour per time step of G)

With `M`: `run_navierstokes3D(; nx=255, nt=NT)` of scripts/NavierStokes3D_multi_gpu.jl on one rank (the same grid; the
variant the weak-scaling runs use), written to tests/golden/jl_reference_config_B_M.json.

Writes tests/golden/jl_reference_config_B.json after every step: PT iteration count, every residual of every
check, SHA-256 of Pr, Vx, Vy, Vz, C (full local arrays, column-major bytes) and a few sample values.
"""
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)

from oracle import jl_run                     # noqa: E402
from tests import jl_cases as J               # noqa: E402

OUT = os.path.join(ROOT, "tests", "golden", "jl_reference_config_B.json")
SAMPLES = [(127, 76, 76), (51, 76, 76), (60, 70, 10), (200, 100, 140), (1, 1, 1), (254, 152, 152)]   # 0-based


def main():
    nt = int(sys.argv[1]) if len(sys.argv) > 1 else 2
    variant = sys.argv[2] if len(sys.argv) > 2 else "G"
    global OUT
    if variant == "M":
        OUT = OUT.replace("config_B.json", "config_B_M.json")
    rec = {"script": "scripts/NavierStokes3D_gpu.jl" if variant == "G" else "scripts/NavierStokes3D_multi_gpu.jl",
           "what": "runme() as shipped, first time steps" if variant == "G" else "run_navierstokes3D(nx=255) on one rank, first time steps",
           "steps": []}
    t0 = time.time()

    def on_step(it, env):
        rec["grid"] = [env["nx"], env["ny"], env["nz"]]
        rec["params"] = {k: env[k] for k in ("dx", "dy", "dz", "dt", "dτ", "damp", "niter", "nchk", "g", "ox", "a2")}
        rec["steps"].append({
            "it": it, "iters": int(env["iter"]), "errs": [float(e) for e in env["err_evo"]],
            "digest": {n: J.digest(env[n]) for n in J.RUN_FIELDS},
            "samples": {n: [float(env[n][i]) for i in SAMPLES] for n in J.RUN_FIELDS},
            "seconds": round(time.time() - t0, 1),
        })
        with open(OUT, "w") as fh:
            json.dump(rec, fh, indent=1, ensure_ascii=False)
        print("step", it, "iterations", int(env["iter"]), "after", round(time.time() - t0), "s", flush=True)

    if variant == "M":
        jl_run.run_M(255, nt, on_step=on_step)
        return
    # nx stays the script's literal 255 (G:44): freeze only nt
    from oracle.jl_interp import JuliaScript
    s = JuliaScript.from_file(jl_run.G_PATH)
    s.frozen = {"nt": nt}
    head = s.find_line(r"function runme\(")
    prefix = (head + 1, s.find_line(r"^\s*if do_save !ispath", head) - 1)
    first = s.find_line(r"^\s*for it = 1:nt", prefix[1])
    loop = (first, s.find_line(r"^\s*if do_vis && it % nvis == 0", first) - 1)
    jl_run._run(s, prefix, loop, {"do_vis": False, "do_save": False}, nt, on_step)


if __name__ == "__main__":
    main()
