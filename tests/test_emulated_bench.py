"""bench.py's native arm, executed on the CPU: the emulated library behind the binding and the handful
of torch.cuda objects the bench uses replaced by stand-ins (events that read the host clock, a
stream wrapper, pin_memory as identity).  Numbers are meaningless here; what is checked is that the
arm runs end to end and prints one JSON line with the driver's contract keys -- the bench line is
the most consequential thing the GPU box produces, and this is the only place its code path runs
without one."""
import json
import sys
import time

import pytest

from tests import emu


class _Event:
    def __init__(self, enable_timing=False):
        self.t = None

    def record(self, stream=None):
        self.t = time.perf_counter()

    def elapsed_time(self, other):
        return (other.t - self.t) * 1e3


class _Stream:
    def __init__(self, ptr, device=None):
        self.ptr = ptr


def test_native_arm_prints_the_contract_line(monkeypatch, capsys):
    import torch
    import bench
    monkeypatch.setattr(torch.cuda, "set_device", lambda d: None)
    monkeypatch.setattr(torch.cuda, "synchronize", lambda *a: None)
    monkeypatch.setattr(torch.cuda, "Event", _Event)
    monkeypatch.setattr(torch.cuda, "ExternalStream", _Stream)
    monkeypatch.setattr(torch.Tensor, "pin_memory", lambda self: self.clone())   # a pinned COPY, like the real one
    monkeypatch.setitem(bench.WORKLOADS, "A", ("M", 31, None, None))             # CPU-sized stand-in for 63x38x38
    monkeypatch.setattr(sys, "argv", ["bench.py", "--workload", "A", "--steps", "1", "--warmup", "1", "--mode", "FAST",
                                      "--no-cpu-baseline"])
    with emu.use_emulated_library():
        bench.main()
    line = json.loads(capsys.readouterr().out.strip().splitlines()[-1])
    for k in ("metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling",
              "vs_baseline", "dtype", "data", "config", "clocks", "e2e", "gpu_launches", "roofline", "parity_check"):
        assert k in line, k
    assert line["metric"] == "T_eff" and line["unit"] == "GB/s" and line["dtype"] == "f64" and line["n_gpus"] == 1
    assert line["config"]["workload"].startswith("A: cylinder flow 31x19x19") and line["config"]["mode"] == "FAST"
    assert line["gpu_launches"] > 0 and line["value"] > 0 and line["steps"] == 1
    assert set(("bound", "achieved", "peak", "unit", "frac", "traffic", "dram_achieved", "dram_frac")) <= set(line["roofline"])
    assert "ptv_kernel" in line["roofline"]["kernel"] and line["roofline"]["pt_iterations_per_launch"] == 2
    assert line["e2e"]["h2d_bytes_per_step"] > 0 and line["e2e"]["d2h_bytes_per_step"] > 0 and line["e2e"]["value"] > 0
    assert line["parity_check"]["pt_iters_identical"] and line["parity_check"]["within_tolerance"]
    assert line["parity_check"]["max_rel_diff"]["Pr"] == 0.0     # FAST equals PARITY in every value seen so far
    assert len(line["pt_iters_per_step"]) == 1 and line["pt_iters_per_step"][0] > 18


def test_warm_up_steps_are_checked_against_the_reference_text(monkeypatch, capsys, tmp_path):
    """Workload B's warm-up steps are compared with the record of the shipped script's text
    (tests/golden/jl_reference_config_B.json).  Here: a CPU-sized stand-in for B (script G at nx = 20) and the
    record built from the G20 run of tests/golden/jl_reference_fixtures.npz; then a falsified record must fail the run."""
    import os

    import numpy as np
    import torch

    import bench
    z = np.load(os.path.join(bench.ROOT, "tests", "golden", "jl_reference_fixtures.npz"))
    m = json.loads(str(z["meta"]))["run"]["G20"]
    # the fixture keeps the final state only: a one-step record would need step 1's digest -> check counts and residuals
    # of both steps and the digest after the second (= the fixture's final state)
    rec = {"grid": [20, 12, 12], "steps": [{"it": 1, "iters": m["iters"][0], "errs": m["errs"][0], "digest": {}},
                                           {"it": 2, "iters": m["iters"][1], "errs": m["errs"][1], "digest": m["digest"]}]}
    path = tmp_path / "record.json"
    monkeypatch.setattr(torch.cuda, "set_device", lambda d: None)
    monkeypatch.setattr(torch.cuda, "synchronize", lambda *a: None)
    monkeypatch.setattr(torch.cuda, "Event", _Event)
    monkeypatch.setattr(torch.cuda, "ExternalStream", _Stream)
    monkeypatch.setattr(torch.Tensor, "pin_memory", lambda self: self.clone())
    monkeypatch.setitem(bench.WORKLOADS, "B", ("G", 20, None, None))
    monkeypatch.setattr(bench, "TEXT_RECORD", str(path))
    monkeypatch.setattr(sys, "argv", ["bench.py", "--steps", "1", "--warmup", "2", "--no-cpu-baseline", "--no-extras"])
    path.write_text(json.dumps(rec))
    with emu.use_emulated_library():
        bench.main()
    line = json.loads(capsys.readouterr().out.strip().splitlines()[-1])
    t = line["parity_check"]["reference_text"]
    assert t == {"against": t["against"], "steps_checked": 2, "pt_iters_identical": True, "residuals_identical": True,
                 "fields_bit_identical": True}
    rec["steps"][1]["errs"][-1] *= 1.0000000001
    path.write_text(json.dumps(rec))
    with emu.use_emulated_library():
        with pytest.raises(SystemExit) as e:
            bench.main()
    assert e.value.code == 3
