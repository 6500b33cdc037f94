#!/bin/bash
mkdir -p gpurun_out/r2c6 gpurun_out/r2ncu && cd "$(dirname "$0")/../.." || exit 1
O=gpurun_out/r2c6
timeout 600 python -m pytest tests/test_gpu_solver.py -m gpu -x -q -k "ptv or fused_iteration" > $O/pytest_ptv.log 2>&1; rc=$?; echo "pytest ptv rc=$rc"; tail -3 $O/pytest_ptv.log
if [ $rc -ne 0 ]; then grep -n "Error\|error\|assert" $O/pytest_ptv.log | head -20; exit 1; fi
timeout 300 python tools/sweep_ptv.py --grids 255x153x153 --modes FAST,FASTEST --old --sets "ptv_k=2,ptv_lb=1,ptv_pxt=16,ptv_bty=16;ptv_k=2,ptv_lb=0,ptv_pxt=16,ptv_bty=16;ptv_k=2,ptv_lb=1,ptv_pxt=32,ptv_bty=8;ptv_k=2,ptv_lb=0,ptv_pxt=32,ptv_bty=8;ptv_k=2,ptv_lb=3,ptv_pxt=32,ptv_bty=16;ptv_k=2,ptv_lb=4,ptv_pxt=32,ptv_bty=16;ptv_k=2,ptv_lb=3,ptv_pxt=16,ptv_bty=32;ptv_k=2,ptv_lb=1,ptv_pxt=16,ptv_bty=16,zchunk=19;ptv_k=2,ptv_lb=1,ptv_pxt=16,ptv_bty=16,zchunk=13;ptv_k=2,ptv_lb=1;ptv_k=3,ptv_lb=0,ptv_pxt=16,ptv_bty=16;ptv_k=3,ptv_lb=3,ptv_pxt=16,ptv_bty=32;ptv_k=3,ptv_lb=3,ptv_pxt=32,ptv_bty=16;ptv_k=1,ptv_lb=1,ptv_pxt=32,ptv_bty=8" > $O/sweep_B.jsonl 2> $O/sweep_B.err; echo "sweep B rc=$?"; cut -c1-235 $O/sweep_B.jsonl
timeout 300 python tools/sweep_ptv.py --grids 511x511x511 --modes FASTEST --iters 48 --reps 2 --sets "ptv_k=2,ptv_lb=1,ptv_pxt=16,ptv_bty=16;ptv_k=2,ptv_lb=1,ptv_pxt=32,ptv_bty=8;ptv_k=2,ptv_lb=3,ptv_pxt=32,ptv_bty=16;ptv_k=3,ptv_lb=0,ptv_pxt=16,ptv_bty=16;ptv_k=3,ptv_lb=3,ptv_pxt=32,ptv_bty=16" > $O/sweep_511.jsonl 2> $O/sweep_511.err; echo "sweep 511 rc=$?"; cut -c1-235 $O/sweep_511.jsonl
TAG=d bash tools/gpu_r2_ncu.sh ptv_k=2,ptv_lb=1,ptv_pxt=16,ptv_bty=16
echo "elapsed ${SECONDS}s"
