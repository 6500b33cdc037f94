#!/bin/bash
mkdir -p gpurun_out/r2c9 && cd "$(dirname "$0")/../.." || exit 1
O=gpurun_out/r2c9
S=""
for k in 2 3; do for z in 16 32 64; do
  S="$S;ptv_k=$k,zchunk=$z;ptv_k=$k,ptv_lb=0,zchunk=$z;ptv_k=$k,ptv_lb=3,ptv_pxt=32,ptv_bty=16,zchunk=$z;ptv_k=$k,ptv_lb=1,ptv_pxt=32,ptv_bty=8,zchunk=$z"
done; done
S=${S#;}
timeout 600 python tools/sweep_ptv.py --grids 511x511x511 --modes FAST,FASTEST --iters 48 --reps 2 --sets "$S" > $O/sweep_511.jsonl 2> $O/sweep_511.err; echo "sweep 511 rc=$?"; cut -c1-200 $O/sweep_511.jsonl
timeout 300 python tools/sweep_ptv.py --grids 1023x511x511 --modes FAST --iters 24 --reps 2 --old --sets "ptv_k=2,zchunk=32;ptv_k=3,ptv_lb=0,zchunk=32;ptv_k=2,ptv_lb=3,ptv_pxt=32,ptv_bty=16,zchunk=32" > $O/sweep_D.jsonl 2> $O/sweep_D.err; echo "sweep D rc=$?"; cut -c1-200 $O/sweep_D.jsonl
echo "elapsed ${SECONDS}s"
