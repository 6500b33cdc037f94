#!/bin/bash
# Round 2, GPU call 4 (one B200): the TMA-staged kernel on hardware -- its parity cases first (bounded waits), then sweeps.
mkdir -p gpurun_out/r2c4 && cd "$(dirname "$0")/../.." || exit 1
O=gpurun_out/r2c4
timeout 600 python -m pytest tests/test_gpu_solver.py -m gpu -x -q -k "ptv or fused_iteration" > $O/pytest_ptv.log 2>&1; rc=$?; echo "pytest ptv rc=$rc"; tail -4 $O/pytest_ptv.log
if [ $rc -ne 0 ]; then grep -n "Error\|error\|assert" $O/pytest_ptv.log | head -20; exit 1; fi
timeout 300 python tools/sweep_ptv.py --grids 255x153x153 --modes FAST,FASTEST --old > $O/sweep_B.jsonl 2> $O/sweep_B.err; echo "sweep B rc=$?"; cut -c1-250 $O/sweep_B.jsonl
timeout 300 python tools/sweep_ptv.py --grids 255x153x153 --modes FAST --sets "ptv_k=2,ptv_lb=1,ptv_pxt=32,ptv_bty=8;ptv_k=2,ptv_lb=1,ptv_pxt=16,ptv_bty=16;ptv_k=2,ptv_lb=1,ptv_pxt=64,ptv_bty=4;ptv_k=2,ptv_lb=4,ptv_pxt=32,ptv_bty=16;ptv_k=2,ptv_lb=4,ptv_pxt=16,ptv_bty=32;ptv_k=2,ptv_lb=1,ptv_ns=3;ptv_k=2,ptv_lb=1,ptv_ns=5;ptv_k=2,ptv_lb=1,ptv_ns=6;ptv_k=2,ptv_lb=1,ptv_tma=0;ptv_k=2,ptv_lb=1,zchunk=12;ptv_k=2,ptv_lb=1,zchunk=19;ptv_k=2,ptv_lb=1,zchunk=25;ptv_k=2,ptv_lb=1,zchunk=38;ptv_k=2,ptv_lb=2,ptv_pxt=32,ptv_bty=8;ptv_k=2,ptv_lb=0,ptv_pxt=32,ptv_bty=8" > $O/sweep_B2.jsonl 2> $O/sweep_B2.err; echo "sweep B2 rc=$?"; cut -c1-250 $O/sweep_B2.jsonl
timeout 300 python tools/sweep_ptv.py --grids 511x511x511 --modes FAST,FASTEST --iters 48 --reps 2 --old > $O/sweep_511.jsonl 2> $O/sweep_511.err; echo "sweep 511 rc=$?"; cut -c1-250 $O/sweep_511.jsonl
echo "elapsed ${SECONDS}s"
