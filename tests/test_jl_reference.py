"""The oracle pinned to the reference's own SOURCE TEXT.

Julia is not in the image, so the reference cannot run; its one golden vector is stale.  What can be
done instead: oracle/jl_interp.py tokenises and parses the two reference scripts as they lie under
/root/reference and evaluates their kernels, boundary-condition functions, parameter blocks, initial
conditions and time loops with numpy -- only the meaning of the package names (ParallelStencil's
FiniteDifferences3D macros and launch ranges, ImplicitGlobalGrid on one rank, Base Julia arithmetic) is
restated there.  tests/golden/make_jl_fixtures.py stored what that execution produces.

* always (the fixtures travel, /root/reference does not): the C oracle against the fixtures, BIT-EXACT --
  78 single launches on seeded random fields and 6 whole runs incl. PT iteration counts and err histories;
* when /root/reference is present: the fixtures re-derived from the text (they cannot drift from it), and
  the line ranges the runs execute checked against the cited ones;
* the interpreter itself on Julia snippets whose value is known (precedence, `2μ`, `^`, `%`, short-circuit,
  divergent `if`, bounds checks).
"""
import json
import math
import os

import numpy as np
import pytest

from oracle import jl_run
from oracle.jl_interp import JlError, JuliaScript, L, lin_range
from tests import jl_cases as J

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
live = pytest.mark.skipif(not jl_run.reference_available(), reason="/root/reference is not on this machine")


@pytest.fixture(scope="module")
def fx():
    z = np.load(os.path.join(GOLD, "jl_reference_fixtures.npz"))
    return z, json.loads(str(z["meta"]))


def same_bits(a, b):
    return a.shape == b.shape and np.array_equal(np.asfortranarray(a).view(np.uint64), np.asfortranarray(b).view(np.uint64))


# ---- the C oracle against what the reference's text computes ----------------------------------------
@pytest.mark.parametrize("case", J.KERNEL_CASES, ids=J.case_id)
def test_oracle_kernel_equals_reference_text(O, fx, case):
    z, meta = fx
    p, f = J.inputs_of(O, case)
    before = {n: a.copy(order="F") for n, a in f.items()}
    J.run_oracle(O, case, p, f)
    cid = J.case_id(case)
    for n in J.OUTPUTS[case[0]]:
        assert J.digest(f[n]) == meta["kernel"][cid][n], f"{cid}: {n} differs from the reference text's result"
        key = f"kernel/{cid}/{n}"
        if key in z.files:
            assert same_bits(f[n], z[key]), key
    for n in f:                                   # and nothing but the outputs is touched
        if n not in J.OUTPUTS[case[0]] and n != "absRp":
            assert same_bits(f[n], before[n]), f"{cid}: {n} modified"


@pytest.mark.parametrize("rc", J.RUN_CASES, ids=lambda rc: rc[0])
def test_oracle_run_equals_reference_text(O, fx, rc):
    z, meta = fx
    fields, iters, errs, p = J.run_case_oracle(O, rc)
    m = meta["run"][rc[0]]
    assert iters == m["iters"]
    assert errs == m["errs"]                       # every residual of every check, to the last bit
    for n in J.RUN_FIELDS:
        assert J.digest(fields[n]) == m["digest"][n], f"{rc[0]}: {n}"
        if rc[0] in J.FULL_ARRAYS:
            assert same_bits(fields[n], z[f"run/{rc[0]}/{n}"])
    # the parameter block (M:290-341 / G:15-61) as the text derives it
    for jl, attr in (("dx", "dx"), ("dy", "dy"), ("dz", "dz"), ("dt", "dt"), ("dτ", "dtau"), ("damp", "damp"),
                     ("niter", "niter"), ("nchk", "nchk"), ("a2", "a2"), ("b2", "b2"), ("g", "g")):
        assert getattr(p, attr) == m["params"][jl], jl


def test_test3D_samples_of_the_text_run(O, fx):
    """test/test3D.jl's own sampling `Pr[inds_x,inds_y,inds_z]` on the M63 run of the text (three steps: the
    first is degenerate): the oracle reproduces the 64 samples bit for bit, and they are NOT the stale literals."""
    z, meta = fx
    with open(os.path.join(GOLD, "test3D_pr_ref.json")) as fh:
        t3 = json.load(fh)
    fields, _, _, _ = J.run_case_oracle(O, next(rc for rc in J.RUN_CASES if rc[0] == "M63"))
    ix, iy, iz = (np.array(t3[k]) - 1 for k in ("inds_x", "inds_y", "inds_z"))
    got = O.interior(fields["Pr"])[np.ix_(ix, iy, iz)]
    assert same_bits(got, z["run/M63/Pr_samples"])
    assert np.abs(got).max() > 0
    assert not np.allclose(got, np.array(t3["Pr_ref"]).transpose(2, 1, 0), rtol=t3["rtol"], atol=0.0)


@pytest.mark.parametrize("case", J.RANK_CASES, ids=lambda c: c[0])
def test_igg_emulation_equals_the_text_on_several_ranks(O, fx, case):
    """the oracle's ImplicitGlobalGrid emulation (VirtualRanks -- the truth of every multi-GPU
    test) against the script's text interpreted by one thread per rank, `update_halo!` exchanging planes at the
    text's ten call sites and `max_g` reducing over the ranks: all 17 local arrays of every rank bit for bit, for
    z-slabs (the product's decomposition) and for x / y-z splits (halo order x -> y -> z, corners)."""
    z, meta = fx
    got = run = J.run_ranks_oracle(O, case)
    want = meta["ranks"][case[0]]
    assert len(got) == len(want) == int(np.prod(case[4]))
    for r, ((f, iters, errs), w) in enumerate(zip(run, want)):
        assert iters == w["iters"] and errs == w["errs"]
        for n in J.RANK_FIELDS:
            assert J.digest(f[n]) == w["digest"][n], f"{case[0]} rank {r}: {n}"


# ---- the fixtures re-derived from the text (build container only) -------------------------------------
@live
@pytest.mark.parametrize("case", J.RANK_CASES, ids=lambda c: c[0])
def test_fixtures_are_what_the_text_computes_ranks(fx, case):
    z, meta = fx
    for (f, iters, errs), w in zip(J.run_ranks_interp(jl_run, case), meta["ranks"][case[0]]):
        assert iters == w["iters"] and errs == w["errs"]
        assert {n: J.digest(f[n]) for n in J.RANK_FIELDS} == w["digest"]


@live
def test_the_shipped_multi_rank_geometry_diverges_in_the_text_too():
    """With its own rule (lz fixed, nz_g growing with the ranks: dz != dx) the script's PT loop diverges on two
    ranks at this size and `floor(Int, .)` in backtrack! throws -- in the text as in Julia.  Recorded, not a gate."""
    with pytest.raises(RuntimeError, match="InexactError"):
        jl_run.run_M_ranks(24, 2, (1, 1, 2))


@live
def test_return_value_of_run_navierstokes3D():
    """M:375-403 + M:528-535 executed as well: the five returned arrays are the interiors (what test3D.jl:6 receives)."""
    env, iters, errs, info = jl_run.run_M(31, 2, returns=True)
    assert info["out_alloc"] == (375, 403) and info["ret"] == (528, 532)
    for n in ("C", "Pr", "Vx", "Vy", "Vz"):
        assert np.array_equal(env[n + "_v"], env[n][1:-1, 1:-1, 1:-1])
    assert env["Vx_v"].shape == (30, 17, 17) and env["Pr_v"].shape == (29, 17, 17)


@pytest.fixture(scope="module")
def scripts():
    return {"M": JuliaScript.from_file(jl_run.M_PATH), "G": JuliaScript.from_file(jl_run.G_PATH)}


@live
def test_fixtures_are_what_the_text_computes_kernels(O, fx, scripts):
    z, meta = fx
    for case in J.KERNEL_CASES:
        p, f = J.inputs_of(O, case)
        J.run_interp(scripts[case[1]], case, p, f)
        for n in J.OUTPUTS[case[0]]:
            assert J.digest(f[n]) == meta["kernel"][J.case_id(case)][n], (J.case_id(case), n)


@live
@pytest.mark.parametrize("rid", ["M31", "G20", "M40rot"])
def test_fixtures_are_what_the_text_computes_runs(fx, rid):
    z, meta = fx
    rc = next(rc for rc in J.RUN_CASES if rc[0] == rid)
    fields, iters, errs, env, info = J.run_case_interp(jl_run, rc)
    assert iters == meta["run"][rid]["iters"] and errs == meta["run"][rid]["errs"]
    for n in J.RUN_FIELDS:
        assert J.digest(fields[n]) == meta["run"][rid]["digest"][n]


@live
def test_the_runs_execute_the_cited_lines(scripts):
    """M:288-373 + M:446-477, G:13-88 + G:119-142 (what DESIGN.md and the oracle cite)."""
    _, _, _, info = jl_run.run_M(9, 1)
    assert info["prefix"] == (288, 374) and info["loop"] == (446, 477)
    _, _, _, info = jl_run.run_G(20, 1)
    assert info["prefix"] == (13, 88) and info["loop"] == (119, 142)
    kinds = {n: d.kind for n, d in scripts["M"].defs.items()}
    assert kinds["update_dPrdτ!"] == "ps_kernel" and kinds["advect!"] == "pi_kernel" and kinds["backtrack!"] == "function"
    assert scripts["M"].defs["update_τ!"].lines == (36, 44) and scripts["G"].defs["set_cylinder!"].lines == (336, 368)
    assert set(scripts["M"].macros) == {"∇V"} == set(scripts["G"].macros)


@live
def test_launch_sequence_of_one_time_step():
    """The order of kernel launches in one step of M, as the text issues them (M:449-476): what ns3d_step fuses."""
    env, iters, _, info = jl_run.run_M(9, 1)
    names = [n for n, _ in info["script"].launches]
    assert names[0] == "set_cylinder!"                                  # M:372
    step = names[1:]
    assert step[:4] == ["update_τ!", "predict_V!", "set_cylinder!", "update_∇V!"]
    per_iter = ["update_dPrdτ!", "update_Pr!", "bc_x!", "bc_y!", "bc_z!", "bc_x_Pr!"]
    assert step[4:4 + len(per_iter)] == per_iter
    tail = ["correct_V!", "set_cylinder!", "bc_x!", "bc_y!", "bc_z!", "bc_x!", "bc_z!", "bc_x!", "bc_y!", "bc_x_Vx!", "advect!"]
    assert step[-len(tail):] == tail
    assert step.count("compute_res!") == len(range(env["nchk"], iters[0] + 1, env["nchk"]))


@live
def test_a_diverged_run_raises_like_julia(scripts):
    """`floor(Int, x)` of a non-finite value is an InexactError in Julia; the interpreter does not paper over it."""
    S = scripts["M"]
    n = (5, 4, 3)
    A, A_o = np.zeros(n, order="F"), np.zeros(n, order="F")
    ix = L(np.array([2])); one = L(np.array([1]))
    with pytest.raises(JlError, match="InexactError"):
        S.call_def(S.defs["backtrack!"], [A, A_o, L(np.array([np.inf])), 0.0, 0.0, 1.0, 1.0, 1.0, 1.0, ix, one, one])


# ---- the interpreter on snippets with known values -------------------------------------------------------
def run_snippet(src, env=None, **frozen):
    s = JuliaScript(src)
    s.frozen = frozen
    env = {} if env is None else env
    s.run_lines(1, src.count("\n") + 1, env)
    return env, s


def test_precedence_and_literal_coefficients():
    env, _ = run_snippet("μ = 3.0\na = 2μ*5.0\nb = 1/2μ\nc = -2.0^2\nd = 2.0^3^2\ne = 7 % 3\nf = -7.5 % 2\n"
                         "g = 1.0/3.0/3.0\nh = 1.0 - 2.0 - 3.0\ni = 2μ^2\nj = (1 + 2)*3\nk = 10/4\nl = 1:3\nm = 2 < 3 && 3 < 2 || true\n"
                         "n = 1/Inf^2*5.0\no = ceil(Int, 63*0.6)\np = !(1 > 2)")
    assert env["a"] == 30.0 and env["b"] == 1 / 6.0 and env["c"] == -4.0 and env["d"] == 512.0
    assert env["e"] == 1 and env["f"] == -1.5                        # rem: sign of the dividend, like C fmod
    assert env["g"] == (1.0 / 3.0) / 3.0 and env["h"] == -4.0 and env["i"] == 18.0 and env["j"] == 9 and env["k"] == 2.5
    assert env["l"] == ("range", 1, 3) and env["m"] is True and env["n"] == 0.0 and env["o"] == 38 and env["p"] is True


def test_tuple_assignment_blocks_and_loops():
    env, _ = run_snippet("a, b = 1.5, 2.5\ns, c = sincos(0*π/6)\nacc = 0\nfor i = 1:10\n  if i % 2 == 0 acc += i end\n"
                         "  if i == 7 break end\nend\nxs = Float64[]; push!(xs, acc/4)\nf(x,y) = (t = x*y; t + 1)\nv = f(2, 3)")
    assert (env["a"], env["b"], env["s"], env["c"]) == (1.5, 2.5, 0.0, 1.0)
    assert env["acc"] == 12 and env["i"] == 7 and env["xs"] == [3.0] and env["v"] == 7


def test_frozen_literals():
    env, _ = run_snippet("nx = 255\nny = ceil(Int, nx*0.6)", nx=20)
    assert env["nx"] == 20 and env["ny"] == 12


def test_linrange_is_lerpi():
    r = lin_range(-0.5, 0.5, 4)
    assert r[0] == -0.5 and r[3] == 0.5 and r[1] == (1 - 1 / 3) * -0.5 + (1 / 3) * 0.5


SRC_KERNELS = '''
@parallel function lap!(B, A, dx)
    @inn(B) = @d2_xi(A)/dx/dx + @d2_yi(A)/dx/dx + @d2_zi(A)/dx/dx
    @all(B) = @all(B) + 1.0
    return
end
@parallel_indices (ix,iy,iz) function pick!(A, B)
    if ix > 1 && ix <= size(B,1) && B[ix-1,iy,iz] > 0.0
        t = B[ix-1,iy,iz]
        if t > 0.5
            A[ix,iy,iz] = t
        else
            A[ix,iy,iz] = -t
        end
    end
    return
end
@parallel_indices (iy,iz) function oob!(A)
    A[end+1, iy, iz] = 0.0
    return
end
function go!(A, B, dx)
    @parallel lap!(B, A, dx)
    @parallel (1:size(A,2),1:size(A,3)) oob!(A)
    return
end
'''


def test_parallel_function_statement_boxes():
    s = JuliaScript(SRC_KERNELS)
    rng = np.random.default_rng(3)
    A = np.asfortranarray(rng.uniform(-1, 1, (6, 5, 4)))
    B = np.zeros((6, 5, 4), order="F")
    s.launch(s.defs["lap!"], [B, A, 0.5])
    want = np.ones_like(B)
    for i in range(1, 5):
        for j in range(1, 4):
            for k in range(1, 3):
                d2 = lambda a, b, c: (a - b) - (b - c)   # noqa: E731
                want[i, j, k] = ((d2(A[i + 1, j, k], A[i, j, k], A[i - 1, j, k]) / 0.5 / 0.5
                                  + d2(A[i, j + 1, k], A[i, j, k], A[i, j - 1, k]) / 0.5 / 0.5)
                                 + d2(A[i, j, k + 1], A[i, j, k], A[i, j, k - 1]) / 0.5 / 0.5) + 1.0
    assert same_bits(B, want)


def test_divergent_if_short_circuit_and_bounds():
    s = JuliaScript(SRC_KERNELS)
    rng = np.random.default_rng(4)
    B = np.asfortranarray(rng.uniform(-1, 1, (5, 4, 3)))
    A = np.full((6, 4, 3), 9.0, order="F")            # the launch box is A's: ix reaches 6 > size(B,1)
    s.launch(s.defs["pick!"], [A, B])
    want = np.full_like(A, 9.0)
    for i in range(1, 5):                             # 0-based ix-1 in 1..4  <->  Julia ix in 2..5
        for j in range(4):
            for k in range(3):
                t = B[i - 1, j, k]
                if t > 0.0:
                    want[i, j, k] = t if t > 0.5 else -t
    assert same_bits(A, want)                         # and B[ix-1] was never read where ix == 1 or ix == 6
    with pytest.raises(JlError, match="out of bounds"):
        s.call_def(s.defs["go!"], [A, np.zeros_like(A), 1.0])


def test_unknown_names_fail_loudly():
    with pytest.raises(JlError, match="unknown name"):
        run_snippet("a = no_such_function(1)")
    assert math.isnan(run_snippet("a = 0.0/0.0")[0]["a"])


# ---- sensitivity: an edited text gives other bits -------------------------------------------------------------
MUTATIONS = [
    # (kernel case whose result must change, text before, text after)
    (("update_dPrdτ!", "M", (13, 9, 8), None), "@d2_xi(Pr)/dx/dx", "@d2_xi(Pr)/(dx*dx)"),          # association of the divisions
    (("update_dPrdτ!", "M", (13, 9, 8), None), "- ρ/dt*@inn(∇V))", "- ρ*@inn(∇V)/dt)"),            # (ρ/dt)*x vs (ρ*x)/dt
    (("update_τ!", "M", (13, 9, 8), None), "@∇V()/3.0)", "@∇V()*(1.0/3.0))"),                       # division vs reciprocal
    (("predict_V!", "G", (13, 9, 8), None), "@d_ya(τyz)/dy - ρ*g)", "@d_ya(τyz)/dy) - dt*g"),      # gravity outside the dt/ρ factor
    (("correct_V!", "M", (13, 9, 8), None), "dt/ρ*@d_xi(Pr)/dx", "dt*@d_xi(Pr)/(ρ*dx)"),
    (("advect!", "M", (13, 9, 8), 3.0), "lerp(a,b,t) = b*t + a*(1-t)", "lerp(a,b,t) = a + (b-a)*t"),
    (("advect!", "M", (13, 9, 8), 3.0), "backtrack!(Vy,Vy_o,vxc,vyc,vzc,dt,dx,dy,dz,ix,iy,iz)\n    end\n    if checkbounds",
     "backtrack!(Vz,Vz_o,vxc,vyc,vzc,dt,dx,dy,dz,ix,iy,iz)\n    end\n    if checkbounds"),      # "fixing" quirk 1 (M:234)
    (("advect!", "M", (13, 9, 8), 0.0), "δx = (δx>0) - (δx%1)", "δx = (δx>=0) - (δx%1)"),          # weight at zero displacement
    (("set_cylinder!", "G", (31, 19, 4), "script"), "xc,yc,zc = xv+dx/2, yv+dx/2, zv+dz/2", "xc,yc,zc = xv+dx/2, yv+dy/2, zv+dz/2"),   # "fixing" G:338
    (("set_cylinder!", "M", (40, 24, 3), "rotated"), "< 1.05", "<= 1.0"),
]


@live
@pytest.mark.parametrize("mut", MUTATIONS, ids=lambda m: f"{m[0][0]}:{m[1][:18]}")
def test_an_edited_text_changes_the_bits(O, fx, mut):
    """The fixtures are sensitive to exactly the things a re-typed restatement gets wrong -- association order,
    division vs reciprocal, the scripts' quirks: each one-line edit of the text changes the stored digest."""
    case, old, new = mut
    z, meta = fx
    with open(jl_run.M_PATH if case[1] == "M" else jl_run.G_PATH, encoding="utf-8") as fh:
        text = fh.read()
    assert old in text
    s = JuliaScript(text.replace(old, new))
    p, f = J.inputs_of(O, case)
    J.run_interp(s, case, p, f)
    changed = [n for n in J.OUTPUTS[case[0]] if J.digest(f[n]) != meta["kernel"][J.case_id(case)][n]]
    assert changed, "the edit went unnoticed"


# ---- the benchmark configuration itself (BASELINE configs[1]) ---------------------------------------------------
def config_B_record():
    with open(os.path.join(GOLD, "jl_reference_config_B.json")) as fh:
        return json.load(fh)


def test_oracle_equals_the_shipped_script_at_config_B(O):
    """`runme()` of scripts/NavierStokes3D_gpu.jl exactly as shipped (255x153x153; only `nt` replaced), first time step,
    executed from its text by tests/golden/make_jl_config_B.py (half an hour of numpy): 3 952 PT iterations, 26 residual
    checks.  The C oracle reproduces the count, every residual and all five fields bit for bit (about half a minute)."""
    rec = config_B_record()
    assert rec["grid"] == [255, 153, 153]
    O.lib().ns3d_oracle_set_num_threads(len(os.sched_getaffinity(0)))
    p = O.params_G(255)
    for jl, attr in (("dx", "dx"), ("dt", "dt"), ("dτ", "dtau"), ("damp", "damp"), ("niter", "niter"), ("nchk", "nchk"), ("g", "g"), ("ox", "ox")):
        assert getattr(p, attr) == rec["params"][jl], jl
    f = O.initial_fields(p)
    it, hist = O.step(p, f)
    s = rec["steps"][0]
    assert it == s["iters"] and hist == s["errs"]
    for n in J.RUN_FIELDS:
        assert J.digest(f[n]) == s["digest"][n], n


@live
def test_oracle_equals_the_text_on_random_ragged_shapes(O, scripts):
    """Beyond the stored fixtures: every kernel of both scripts on 8 random shapes in [3, 14]^3 (seeded), the reference's
    text interpreted live against the C oracle, bit for bit."""
    rng = np.random.default_rng(20261018)
    shapes = [tuple(int(v) for v in rng.integers(3, 15, size=3)) for _ in range(8)]
    kernels = ["update_τ!", "predict_V!", "update_∇V!", "update_dPrdτ!", "update_Pr!", "compute_res!", "correct_V!",
               "set_bc_Vel!", "set_bc_Pr!", "advect!", "set_cylinder!"]
    n = 0
    for v in ("M", "G"):
        for g in shapes:
            for k in kernels:
                case = (k, v, g, 2.0 if k == "advect!" else ("script" if k == "set_cylinder!" else None))
                p, f = J.inputs_of(O, case)
                f2 = {a: b.copy(order="F") for a, b in f.items()}
                J.run_oracle(O, case, p, f)
                J.run_interp(scripts[v], case, p, f2)
                for name in f:
                    if name != "absRp":
                        assert same_bits(f[name], f2[name]), (J.case_id(case), name)
                n += 1
    assert n == 2 * 8 * 11


@live
def test_save_path_records_are_what_the_text_writes(fx):
    """The WHOLE bodies of the two run functions (M:288-535, G:13-172; plotting parsed, never reached) with do_save=true."""
    z, meta = fx
    rec = J.save_path_records(jl_run)
    assert rec == meta["save"]
    assert rec["M31"]["lines"] == [288, 535]
    assert sorted(rec["M31"]["files"])[:2] == ["out_C_v_0000.bin", "out_C_v_0001.bin"] and len(rec["M31"]["files"]) == 20
    assert [r["file"] for r in rec["G20"]] == ["out_save/step_0.mat", "out_save/step_1.mat", "out_save/step_2.mat"]
    assert rec["G20"][0]["keys"] == ["C", "Pr", "Vx", "Vy", "dx", "dy", "dz"]            # G:89: "Vy" twice, no "Vz"
    assert rec["G20"][0]["digest"]["Vy"] != rec["G20"][1]["digest"]["Vy"] and "Vz" in rec["G20"][1]["keys"]


@live
def test_whole_function_body_equals_the_stepwise_execution():
    """The fixtures are produced by running the parameter block and then the loop body step by step (to record every
    step's counts and residuals); executing the function body in ONE piece, `for it = 1:nt` and `return` included,
    gives the same return value."""
    ret, env, lines = jl_run.run_M_whole(31, 3)
    env2, iters, errs, info = jl_run.run_M(31, 3, returns=True)
    assert lines == (288, 535)
    for a, n in zip(ret, ("C", "Pr", "Vx", "Vy", "Vz")):
        assert same_bits(a, env2[n + "_v"]), n
    env_g, s, lines_g = jl_run.run_G_whole(20, 2)
    env_g2, _, _, _ = jl_run.run_G(20, 2)
    assert lines_g == (13, 172)
    for n in ("Pr", "Vx", "Vy", "Vz", "C"):
        assert same_bits(env_g[n], env_g2[n]), n


@live
def test_projection_identity_holds_under_the_restated_macro_semantics(O, scripts, monkeypatch):
    """An internal check of the one thing that IS restated (ParallelStencil's finite-difference macros): Chorin's projection
    needs discrete div(grad) == the discrete Laplacian of the PT residual.  With the scripts' own kernels from the text --
    `correct_V!` on V = 0 (`@d_xi`), `update_∇V!` (`@d_xa`, `@all`), `compute_res!` with zero divergence (`@d2_xi`, `@inn`) --
    `∇V_inn == -(dt/ρ)·Rp` must hold to rounding; with a macro restated one cell off it fails at O(1)."""
    S = scripts["M"]
    p = O.params_M(11, 8, 7)
    rng = np.random.default_rng(5)

    def run():
        f = O.alloc_fields(p)
        f["Pr"][...] = rng.uniform(-1, 1, f["Pr"].shape)
        S.launch(S.defs["correct_V!"], [f["Vx"], f["Vy"], f["Vz"], f["Pr"], p.dt, p.rho, p.dx, p.dy, p.dz])
        S.launch(S.defs["update_∇V!"], [f["divV"], f["Vx"], f["Vy"], f["Vz"], p.dx, p.dy, p.dz])
        div = f["divV"].copy()
        f["divV"][...] = 0.0
        S.launch(S.defs["compute_res!"], [f["Rp"], f["Pr"], f["divV"], p.rho, p.dt, p.dx, p.dy, p.dz])
        lhs, rhs = div[1:-1, 1:-1, 1:-1], -(p.dt / p.rho) * f["Rp"]
        return np.abs(lhs - rhs).max() / np.abs(rhs).max()

    assert run() < 1e-12
    orig = S.macrocall

    def off_by_one(name, args, env, ps):          # @d_xi without the inner offset in y and z
        if name == "d_xi" and ps is not None:
            A = S.ev(args[0], env)
            n = ps
            return A[1:1 + n[0], 0:n[1], 0:n[2]] - A[0:n[0], 0:n[1], 0:n[2]]
        return orig(name, args, env, ps)
    monkeypatch.setattr(S, "macrocall", off_by_one)
    assert run() > 0.1


@live
def test_predictor_is_second_order_at_the_scripts_own_node_positions(O, scripts, monkeypatch):
    """A second check of the restated macros, anchored to the reference's text: `set_cylinder!` (M:250-279) says where the
    unknowns live -- `Vx[ix,iy,iz]` at (xv, yc) with `yc = yco_g + (iy-1)*dy`.  For Vx = sin(2π(yc + ly/2)/ly), Vy = Vz = 0 the
    predictor's increment (`update_τ!` then `predict_V!`, from the text) must equal dt·μ/ρ·Vx'' at those positions to SECOND
    order; with `@d_ya` restated one cell off it drops to first order."""
    S = scripts["M"]

    def error(ny):
        p = O.params_M(5, ny, 5)
        f = O.alloc_fields(p)
        yc = p.yco_g + np.arange(ny) * p.dy
        k = 2 * np.pi / p.ly
        f["Vx"][...] = np.sin(k * (yc + p.ly / 2))[None, :, None]
        before = f["Vx"].copy()
        S.launch(S.defs["update_τ!"], [f[n] for n in ("txx", "tyy", "tzz", "txy", "txz", "tyz", "Vx", "Vy", "Vz")] + [p.mu, p.dx, p.dy, p.dz])
        S.launch(S.defs["predict_V!"], [f[n] for n in ("Vx", "Vy", "Vz", "txx", "tyy", "tzz", "txy", "txz", "tyz")]
                 + [p.rho, 0.0, p.dt, p.dx, p.dy, p.dz])
        got = (f["Vx"] - before)[1:-1, 1:-1, 1:-1]
        want = (p.dt * p.mu / p.rho * -k * k * np.sin(k * (yc + p.ly / 2)))[None, 1:-1, None]
        return np.abs(got - want).max() / np.abs(want).max()

    e = [error(n) for n in (16, 32, 64)]
    assert 3.6 < e[0] / e[1] < 4.4 and 3.8 < e[1] / e[2] < 4.2, e
    orig = S.macrocall

    def off_by_one(name, args, env, ps):          # @d_ya read one cell lower
        if name == "d_ya" and ps is not None and args[0] == ("id", "τxy"):
            A = S.ev(args[0], env)
            n = ps
            lo = np.concatenate([A[0:n[0], 0:1, 0:n[2]], A[0:n[0], 0:n[1] - 1, 0:n[2]]], axis=1)
            return A[0:n[0], 0:n[1], 0:n[2]] - lo
        return orig(name, args, env, ps)
    monkeypatch.setattr(S, "macrocall", off_by_one)
    e = [error(n) for n in (16, 32, 64)]
    assert e[1] / e[2] < 2.6, e


@live
def test_two_ranks_of_the_text_reproduce_its_single_rank_run():
    """A check of the restated ImplicitGlobalGrid semantics (overlap 2, which planes `update_halo!` sends, `z_g`, `nz_g`):
    the script's text on two z-ranks against the SAME text on one rank holding the global grid (nz = nz_g = 28): identical
    PT iteration counts, every owned value equal to 1e-13 of the field's scale (not bit-equal: `backtrack!` works in local indices and the
    max-reduction visits the residuals in another order)."""
    lit = {"nz": 15, "lz_lx": 28 / 24}
    ranks = jl_run.run_M_ranks(24, 2, (1, 1, 2), literals=lit)
    env1, iters1, errs1, _ = jl_run.run_M(24, 2, literals={"nz": 28, "lz_lx": 28 / 24})
    assert ranks[0][1] == ranks[1][1] == iters1
    for name, scale in (("Pr", env1["psc"]), ("Vx", env1["vin"]), ("Vy", env1["vin"]), ("C", 1.0)):
        g = env1[name]
        lo, hi = ranks[0][0][name], ranks[1][0][name]
        assert np.abs(lo[:, :, :14] - g[:, :, :14]).max() / scale < 1e-13, name        # rank 0 owns planes 1..14
        assert np.abs(hi[:, :, 1:] - g[:, :, 14:]).max() / scale < 1e-13, name         # rank 1 owns global planes 15..28


@live
def test_the_stale_vector_is_copied_faithfully_from_the_references_test():
    """tests/golden/test3D_pr_ref.json against test/test3D.jl itself: the 64 literals in file order ([z][y][x]), the three
    index rows and the call (`nx=63, nt=1`) -- so that the recorded expected mismatch is a statement about the
    reference's own test, not about a hand copy."""
    import re
    with open(os.path.join(jl_run.REFERENCE_ROOT, "test", "test3D.jl"), encoding="utf-8") as fh:
        t = fh.read()
    with open(os.path.join(GOLD, "test3D_pr_ref.json")) as fh:
        j = json.load(fh)
    block = t[t.index("Pr_ref = ["):t.index("# run reference tests")]
    nums = [float(x) for x in re.findall(r"-?\d+\.\d+(?:e-?\d+)?", block)]
    assert len(nums) == 64 and np.array_equal(np.array(j["Pr_ref"]).ravel(), np.array(nums))
    for k in ("inds_x", "inds_y", "inds_z"):
        row = re.search(k + r"\s*=\s*\[([^\]]*)\]", t).group(1)
        assert [int(v) for v in row.split()] == j[k]
    assert "run_navierstokes3D(do_vis=false, do_save=false, do_print=true, nx=63, nt=1)" in t
    assert j["rtol"] == float(np.sqrt(np.finfo(np.float64).eps))            # Julia's default for `≈`
