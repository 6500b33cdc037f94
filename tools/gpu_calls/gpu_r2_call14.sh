#!/bin/bash
# flow kernel without mbarrier re-initialisation + early claim: the cases that faulted, the whole GPU suite, sweeps, bench
mkdir -p gpurun_out/r2c14 && cd "$(dirname "$0")/../.." || exit 1
O=gpurun_out/r2c14
run() { timeout 120 python tools/debug_ptv_case.py "$@" 2>&1 | tail -3; }
{
run M 63x38x38 0 "ptv_k=2,ptv_lb=0" 40
run M 63x38x38 7 "ptv_k=2,ptv_lb=0" 40 41
run G 70x47x41 2 "ptv_k=2,ptv_lb=0,ptv_pxt=5,ptv_bty=6" 40
} > $O/cases.log 2>&1
cat $O/cases.log
if grep -q FAULT $O/cases.log; then echo "still faulting"; exit 1; fi
timeout 900 python -m pytest tests -m gpu -q -x > $O/pytest_all.log 2>&1; echo "pytest all rc=$?"; tail -4 $O/pytest_all.log
S="ptv_flow=0;ptv_flow=1;ptv_flow=1,zchunk=13;ptv_flow=1,zchunk=16;ptv_flow=1,zchunk=19;ptv_flow=1,zchunk=22;ptv_flow=1,zchunk=26;ptv_flow=1,ptv_k=3,zchunk=26"
timeout 300 python tools/sweep_ptv.py --grids 255x153x153 --modes FAST,FASTEST --iters 152 --sets "$S" > $O/sweep_B.jsonl 2> $O/sweep_B.err; echo "sweep B rc=$?"; cut -c1-200 $O/sweep_B.jsonl
S="ptv_flow=0,zchunk=32;ptv_flow=1,zchunk=32;ptv_flow=1,zchunk=64;ptv_flow=1,ptv_k=3,zchunk=64"
timeout 300 python tools/sweep_ptv.py --grids 511x511x511 --modes FAST,FASTEST --iters 48 --reps 2 --sets "$S" > $O/sweep_511.jsonl 2> $O/sweep_511.err; echo "sweep 511 rc=$?"; cut -c1-200 $O/sweep_511.jsonl
timeout 400 python bench.py --no-extras --no-cpu-baseline > $O/bench_B.json 2> $O/bench_B.err; echo "bench B rc=$?"; cut -c1-300 $O/bench_B.json; tail -3 $O/bench_B.err
echo "elapsed ${SECONDS}s"
