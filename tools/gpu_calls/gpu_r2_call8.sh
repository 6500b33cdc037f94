#!/bin/bash
mkdir -p gpurun_out/r2c8 gpurun_out/r2ncu && cd "$(dirname "$0")/../.." || exit 1
O=gpurun_out/r2c8
timeout 900 python -m pytest tests -m gpu -x -q > $O/pytest.log 2>&1; echo "pytest rc=$?"; tail -3 $O/pytest.log
timeout 300 python tools/sweep_ptv.py --grids 511x511x511 --modes FAST --iters 48 --reps 2 --old --sets "ptv_k=2;ptv_k=2,ptv_lb=0;ptv_k=3;ptv_k=3,ptv_lb=1;ptv_k=2,ptv_lb=3,ptv_pxt=32,ptv_bty=16;ptv_k=2,zchunk=64;ptv_k=3,zchunk=64;ptv_k=3,zchunk=170" > $O/sweep_511_fast.jsonl 2> $O/sweep_511_fast.err; echo "sweep 511 rc=$?"; cut -c1-235 $O/sweep_511_fast.jsonl
timeout 300 python tools/sweep_ptv.py --grids 255x153x153 --modes FAST,FASTEST --sets "ptv_k=2;ptv_k=2,zchunk=8;ptv_k=2,zchunk=10;ptv_k=2,zchunk=16;ptv_k=3;ptv_k=3,zchunk=16" > $O/sweep_B.jsonl 2> $O/sweep_B.err; echo "sweep B rc=$?"; cut -c1-235 $O/sweep_B.jsonl
timeout 400 python bench.py --no-extras > $O/bench_B.json 2> $O/bench_B.err; echo "bench B rc=$?"; cut -c1-300 $O/bench_B.json
timeout 200 python bench.py --workload C --pt-only 1000 --steps 3 --no-cpu-baseline > $O/bench_C.json 2> $O/bench_C.err; echo "bench C rc=$?"; cut -c1-400 $O/bench_C.json
MODE=FAST TAG=e bash tools/gpu_r2_ncu.sh ptv_k=2
echo "elapsed ${SECONDS}s"
