#!/bin/bash
mkdir -p gpurun_out/r2c18 gpurun_out/r2ncu && cd "$(dirname "$0")/../.." || exit 1
O=gpurun_out/r2c18
timeout 300 python -m pytest tests/test_gpu_solver.py -m gpu -x -q -k "ptv_kernel" > $O/pytest_ptv.log 2>&1; echo "pytest ptv rc=$?"; tail -2 $O/pytest_ptv.log
S="ptv_flow=0;ptv_flow=1;ptv_flow=1,zchunk=10;ptv_flow=1,zchunk=16;ptv_flow=1,zchunk=19;ptv_flow=1,ptv_lb=0,zchunk=19"
timeout 300 python tools/sweep_ptv.py --grids 255x153x153 --modes FAST,FASTEST --iters 152 --sets "$S" > $O/sweep_B.jsonl 2> $O/sweep_B.err; echo "sweep B rc=$?"; cut -c1-200 $O/sweep_B.jsonl
S="ptv_flow=0,zchunk=32;ptv_flow=1,zchunk=32;ptv_flow=1,zchunk=64"
timeout 300 python tools/sweep_ptv.py --grids 511x511x511 --modes FAST --iters 48 --reps 2 --sets "$S" > $O/sweep_511.jsonl 2> $O/sweep_511.err; echo "sweep 511 rc=$?"; cut -c1-200 $O/sweep_511.jsonl
timeout 200 ncu --set full --clock-control none --import-source on -k regex:ptv_flow_kernel -c 1 -f -o gpurun_out/r2ncu/flow_vX_B_FAST python tools/profile_pt.py 255x153x153 FAST 0 1 ptv_k=2 152 > $O/ncu_flow.log 2>&1; echo "ncu rc=$?"; tail -1 $O/ncu_flow.log
echo "elapsed ${SECONDS}s"
