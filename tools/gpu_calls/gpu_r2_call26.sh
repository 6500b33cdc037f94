#!/bin/bash
# last sanity of the round on the tree as committed: parity subset + the default bench line
mkdir -p gpurun_out/r2c26 && cd "$(dirname "$0")/../.." || exit 1
O=gpurun_out/r2c26
timeout 200 python -m pytest tests/test_gpu_solver.py tests/test_gpu_zz_output.py -m gpu -x -q -k "time_steps or config_B or step_groups or default_tiles or pt_solve" > $O/pytest_subset.log 2>&1; echo "pytest rc=$?"; tail -2 $O/pytest_subset.log
timeout 300 python bench.py > $O/bench_B.json 2> $O/bench_B.err; echo "bench B rc=$?"; cut -c1-200 $O/bench_B.json; tail -2 $O/bench_B.err
python __graft_entry__.py --smoke 2>&1 | tail -2
echo "elapsed ${SECONDS}s"
