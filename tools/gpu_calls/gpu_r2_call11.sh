#!/bin/bash
# the persistent launch (ptv_flow_kernel) on hardware: parity, then flow vs one-launch-per-pass over chunk lengths
mkdir -p gpurun_out/r2c11 && cd "$(dirname "$0")/../.." || exit 1
O=gpurun_out/r2c11
timeout 600 python -m pytest tests/test_gpu_solver.py -m gpu -x -q > $O/pytest_solver.log 2>&1; rc=$?; echo "pytest solver rc=$rc"; tail -3 $O/pytest_solver.log
if [ $rc -ne 0 ]; then grep -n "Error\|error\|assert" $O/pytest_solver.log | head -20; fi
S="ptv_flow=0;ptv_flow=1;ptv_flow=1,zchunk=10;ptv_flow=1,zchunk=13;ptv_flow=1,zchunk=16;ptv_flow=1,zchunk=19;ptv_flow=1,zchunk=26;ptv_flow=1,zchunk=38;ptv_flow=0,serpentine=0;ptv_flow=1,ptv_k=3,zchunk=16;ptv_flow=1,ptv_k=3,zchunk=26;ptv_flow=1,ptv_lb=0,zchunk=19"
timeout 300 python tools/sweep_ptv.py --grids 255x153x153 --modes FAST,FASTEST --iters 152 --sets "$S" > $O/sweep_B.jsonl 2> $O/sweep_B.err; echo "sweep B rc=$?"; cut -c1-260 $O/sweep_B.jsonl
S="ptv_flow=0,zchunk=32;ptv_flow=1;ptv_flow=1,zchunk=32;ptv_flow=1,zchunk=64;ptv_flow=1,ptv_k=3,zchunk=32;ptv_flow=1,ptv_k=3,zchunk=64"
timeout 300 python tools/sweep_ptv.py --grids 511x511x511 --modes FAST,FASTEST --iters 48 --reps 2 --sets "$S" > $O/sweep_511.jsonl 2> $O/sweep_511.err; echo "sweep 511 rc=$?"; cut -c1-260 $O/sweep_511.jsonl
timeout 400 python bench.py --no-extras --no-cpu-baseline > $O/bench_B.json 2> $O/bench_B.err; echo "bench B rc=$?"; cut -c1-300 $O/bench_B.json; tail -3 $O/bench_B.err
echo "elapsed ${SECONDS}s"
