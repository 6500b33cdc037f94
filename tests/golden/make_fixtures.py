"""Regenerates tests/golden/oracle_fixtures.npz from the CPU oracle (run from the repo root:
``python tests/golden/make_fixtures.py``).  The reference itself cannot run in this container
(Julia absent), so these are ORACLE outputs: they freeze the oracle's behaviour (regression
protection) and give the GPU tests sampled values to compare with even where the oracle is not
rebuilt.  What pins the oracle to the reference is listed in oracle/ns3d_oracle.c."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import oracle as O  # noqa: E402

out = {}
for variant, nx, nt in (("M", 40, 4), ("G", 40, 2), ("M", 63, 3)):
    p = O.params_M(nx) if variant == "M" else O.params_G(nx)
    f, iters, errs = O.run(p, nt)
    key = f"{variant}{nx}"
    out[key + "_iters"] = np.array(iters)
    out[key + "_errs"] = np.array([e[-1] for e in errs])
    rng = np.random.default_rng(1234)
    for name in ("Pr", "Vx", "Vy", "Vz", "C"):
        a = f[name]
        idx = rng.integers(0, a.size, size=64)
        out[f"{key}_{name}_idx"] = idx
        out[f"{key}_{name}_val"] = a.ravel(order="F")[idx]
        out[f"{key}_{name}_sum"] = np.array(a.sum())
np.savez(os.path.join(ROOT, "tests", "golden", "oracle_fixtures.npz"), **out)
print("written", sorted(out)[:6], "...")
