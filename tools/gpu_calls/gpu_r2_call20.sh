#!/bin/bash
# z-bands: passes as bands of launches that overlap consecutive passes (events / graph edges)
mkdir -p gpurun_out/r2c20 && cd "$(dirname "$0")/../.." || exit 1
O=gpurun_out/r2c20
timeout 300 python -m pytest tests/test_gpu_solver.py -m gpu -x -q -k "ptv_kernel or time_steps or config_B" > $O/pytest_ptv.log 2>&1; echo "pytest rc=$?"; tail -2 $O/pytest_ptv.log
S="ptv_bands=0;ptv_bands=3;ptv_bands=4;ptv_bands=6;ptv_bands=8;ptv_bands=4,zchunk=8;ptv_bands=4,zchunk=13;ptv_bands=4,zchunk=19;ptv_bands=8,zchunk=19;ptv_bands=6,zchunk=13"
timeout 300 python tools/sweep_ptv.py --grids 255x153x153 --modes FAST,FASTEST --iters 152 --sets "$S" > $O/sweep_B.jsonl 2> $O/sweep_B.err; echo "sweep B rc=$?"; cut -c1-200 $O/sweep_B.jsonl
S="ptv_bands=0,zchunk=32;ptv_bands=4,zchunk=32;ptv_bands=8,zchunk=32"
timeout 300 python tools/sweep_ptv.py --grids 511x511x511 --modes FAST --iters 48 --reps 2 --sets "$S" > $O/sweep_511.jsonl 2> $O/sweep_511.err; echo "sweep 511 rc=$?"; cut -c1-200 $O/sweep_511.jsonl
echo "elapsed ${SECONDS}s"
