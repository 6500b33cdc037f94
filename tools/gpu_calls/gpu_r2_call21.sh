#!/bin/bash
mkdir -p gpurun_out/r2c21 && cd "$(dirname "$0")/../.." || exit 1
O=gpurun_out/r2c21
S="ptv_bands=8,zchunk=19;ptv_bands=8,zchunk=16;ptv_bands=8,zchunk=22;ptv_bands=6,zchunk=26;ptv_bands=12,zchunk=13;ptv_bands=16,zchunk=10;ptv_bands=14,zchunk=11;ptv_bands=10,zchunk=16;ptv_bands=16,zchunk=8;ptv_bands=8,zchunk=19,ptv_k=3;ptv_bands=6,zchunk=26,ptv_k=3"
timeout 300 python tools/sweep_ptv.py --grids 255x153x153 --modes FAST,FASTEST --iters 152 --sets "$S" > $O/sweep_B.jsonl 2> $O/sweep_B.err; echo "sweep B rc=$?"; cut -c1-150 $O/sweep_B.jsonl
S="ptv_bands=0;ptv_bands=8;ptv_bands=8,zchunk=8;ptv_bands=6,zchunk=10"
timeout 300 python tools/sweep_ptv.py --grids 127x77x77 --modes FAST --iters 120 --sets "$S" > $O/sweep_127.jsonl 2> $O/sweep_127.err; echo "sweep 127 rc=$?"; cut -c1-150 $O/sweep_127.jsonl
echo "elapsed ${SECONDS}s"
