#!/bin/bash
# Round 2, GPU call 3 (one B200): the new pitched-layout kernel on hardware -- parity suite, then sweeps.
mkdir -p gpurun_out/r2c3 && cd "$(dirname "$0")/../.." || exit 1
O=gpurun_out/r2c3
timeout 900 python -m pytest tests -m gpu -x -q > $O/pytest.log 2>&1; echo "pytest rc=$?"; tail -3 $O/pytest.log
timeout 300 python tools/sweep_ptv.py --grids 255x153x153 --modes FAST,FASTEST --old > $O/sweep_B.jsonl 2> $O/sweep_B.err; echo "sweep B rc=$?"; cut -c1-230 $O/sweep_B.jsonl
timeout 300 python tools/sweep_ptv.py --grids 255x153x153 --modes FAST --sets "ptv_k=2,ptv_ry=2,ptv_lb=1,ptv_pxt=128,ptv_bty=3;ptv_k=2,ptv_ry=2,ptv_lb=1,ptv_pxt=65,ptv_bty=5;ptv_k=2,ptv_ry=2,ptv_lb=2,ptv_pxt=65,ptv_bty=3;ptv_k=2,ptv_ry=1,ptv_lb=3,ptv_pxt=32,ptv_bty=8;ptv_k=2,ptv_ry=1,ptv_lb=3,ptv_pxt=65,ptv_bty=3;ptv_k=2,ptv_ry=1,ptv_lb=0,ptv_pxt=65,ptv_bty=7;ptv_k=2,ptv_ry=1,ptv_lb=0,ptv_pxt=128,ptv_bty=4;ptv_k=2,ptv_ry=2,ptv_lb=1,zchunk=12;ptv_k=2,ptv_ry=2,ptv_lb=1,zchunk=19;ptv_k=2,ptv_ry=2,ptv_lb=1,zchunk=76;ptv_k=2,ptv_ry=1,ptv_lb=3,zchunk=12;ptv_k=2,ptv_ry=1,ptv_lb=3,zchunk=19;ptv_k=2,ptv_ry=1,ptv_lb=3,zchunk=38" > $O/sweep_B2.jsonl 2> $O/sweep_B2.err; echo "sweep B2 rc=$?"; cut -c1-230 $O/sweep_B2.jsonl
timeout 300 python tools/sweep_ptv.py --grids 511x511x511 --modes FAST,FASTEST --iters 48 --reps 2 --old > $O/sweep_511.jsonl 2> $O/sweep_511.err; echo "sweep 511 rc=$?"; cut -c1-230 $O/sweep_511.jsonl
timeout 300 python bench.py --no-extras > $O/bench_B.json 2> $O/bench_B.err; echo "bench B rc=$?"; cut -c1-300 $O/bench_B.json
echo "elapsed ${SECONDS}s"
