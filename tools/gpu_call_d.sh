#!/bin/bash
# Round-1 GPU call D: slim kernel with stage 2 ahead of the barrier; tile height / chunk / prefetch sweep, one ncu capture.
mkdir -p gpurun_out
timeout 150 python tools/sweep_pt.py --grids 255x153x153 --modes FASTEST --zchunks 12,16,19 --tb2ty 8,16 \
    --sets "tb2_np=0,tb2_pf=0;tb2_np=1,tb2_pf=0;tb2_np=1,tb2_pf=1;tb2_np=1,tb2_pf=2" \
    > gpurun_out/d_sweep_B.jsonl 2> gpurun_out/d_sweep_B.err
timeout 120 python tools/sweep_pt.py --grids 511x511x511 --modes FASTEST --zchunks 32,64 --tb2ty 8,16 --iters 60 --reps 2 \
    --sets "tb2_np=1,tb2_pf=1;tb2_np=1,tb2_pf=2" > gpurun_out/d_sweep_511.jsonl 2> gpurun_out/d_sweep_511.err
timeout 90 ncu --set full --clock-control none --import-source on -k regex:pt_tb2s -c 2 -f -o gpurun_out/pt_tb2s_ty8_np1_pf1_B_fastest \
    python tools/profile_pt.py 255x153x153 FASTEST 0 1 tb2_ty=8,tb2_np=1,tb2_pf=1 > gpurun_out/d_ncu1.log 2>&1
for f in gpurun_out/d_sweep_B.jsonl gpurun_out/d_sweep_511.jsonl; do python -c "
import sys, json
rows=[json.loads(l) for l in open('$f') if l.strip()]
for r in rows: print(r['tb2_ty'], r['zchunk'], r['opts'], r['us_per_iter'], r['T_eff_GBs'])
"; done
