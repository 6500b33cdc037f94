"""Generates tests/golden/jl_reference_fixtures.npz by EXECUTING THE REFERENCE'S SOURCE TEXT.

    python tests/golden/make_jl_fixtures.py          (needs /root/reference; run in the build container)

oracle/jl_interp.py parses scripts/NavierStokes3D_multi_gpu.jl and scripts/NavierStokes3D_gpu.jl as they
lie under /root/reference and evaluates their kernels, boundary-condition functions, parameter blocks,
initial conditions and time loops with numpy (package semantics restated there: ParallelStencil's
FiniteDifferences3D macros and launch ranges, ImplicitGlobalGrid on one rank, Base Julia arithmetic).
Nothing of the C oracle or of the CUDA library takes part in producing these numbers.

Stored: for every kernel case of tests/jl_cases.py the SHA-256 of each output array (bit-exactness is the
bar) and, for the small grids, the arrays themselves; for the whole runs the PT iteration counts, the err
history, digests of the final Pr,Vx,Vy,Vz,C, the full arrays of the small runs, and -- for the M63 run, the
size of test/test3D.jl -- the 64 samples `Pr[inds_x,inds_y,inds_z]` in that test's own layout; for the
multi-rank cases (one interpreter thread per ImplicitGlobalGrid rank, `update_halo!` at the text's call sites)
the digests of all 17 local arrays of every rank; for the save path (the WHOLE bodies of the two run functions executed
with do_save=true) the SHA-256 of every `out_save/out_<A>_v_%04d.bin` frame and the keys / digests of every `.mat` Dict.
"""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)

from oracle import jl_run, oracle as O                      # noqa: E402
from oracle.jl_interp import JuliaScript                    # noqa: E402
from tests import jl_cases as J                             # noqa: E402

OUT = os.path.join(ROOT, "tests", "golden", "jl_reference_fixtures.npz")
SMALL = 7 * 6 * 5 + 1


def main():
    scripts = {"M": JuliaScript.from_file(jl_run.M_PATH), "G": JuliaScript.from_file(jl_run.G_PATH)}
    out, meta = {}, {"kernel": {}, "run": {}, "lines": {}, "ranks": {}}
    for case in J.KERNEL_CASES:
        p, f = J.inputs_of(O, case)                         # O only supplies shapes and the parameter block
        J.run_interp(scripts[case[1]], case, p, f)
        cid = J.case_id(case)
        meta["kernel"][cid] = {n: J.digest(f[n]) for n in J.OUTPUTS[case[0]]}
        if np.prod(case[2]) < SMALL:
            for n in J.OUTPUTS[case[0]]:
                out[f"kernel/{cid}/{n}"] = f[n]
    with open(os.path.join(ROOT, "tests", "golden", "test3D_pr_ref.json")) as fh:
        t3 = json.load(fh)
    for rc in J.RUN_CASES:
        fields, iters, errs, env, info = J.run_case_interp(jl_run, rc)
        rid = rc[0]
        meta["run"][rid] = {"iters": iters, "errs": errs, "digest": {n: J.digest(fields[n]) for n in J.RUN_FIELDS},
                            "params": {k: env[k] for k in ("dx", "dy", "dz", "dt", "dτ", "damp", "niter", "nchk", "a2", "b2", "g")}}
        meta["lines"][rc[1]] = {"prefix": list(info["prefix"]), "loop": list(info["loop"])}
        if rid in J.FULL_ARRAYS:
            for n in J.RUN_FIELDS:
                out[f"run/{rid}/{n}"] = fields[n]
        if rid == "M63":
            ix, iy, iz = (np.array(t3[k]) - 1 for k in ("inds_x", "inds_y", "inds_z"))
            pr_v = fields["Pr"][1:-1, 1:-1, 1:-1]
            out["run/M63/Pr_samples"] = pr_v[np.ix_(ix, iy, iz)]
        print(rid, iters)
    for case in J.RANK_CASES:
        res = J.run_ranks_interp(jl_run, case)
        meta["ranks"][case[0]] = [{"iters": iters, "errs": errs, "digest": {n: J.digest(f[n]) for n in J.RANK_FIELDS}}
                                  for f, iters, errs in res]
        print(case[0], case[4], res[0][1])
    meta["save"] = J.save_path_records(jl_run)
    print("save path:", len(meta["save"]["M31"]["files"]), "frames,", len(meta["save"]["G20"]), ".mat dumps")
    out["meta"] = np.array(json.dumps(meta, ensure_ascii=False))
    np.savez_compressed(OUT, **out)
    print("wrote", OUT, os.path.getsize(OUT), "bytes;", len(J.KERNEL_CASES), "kernel cases,", len(J.RUN_CASES), "runs")


if __name__ == "__main__":
    main()
