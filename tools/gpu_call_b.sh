#!/bin/bash
# Round-1 GPU call B: sweep of the slim two-iteration kernel's variants (graph cache now keyed on the options), ncu captures.
mkdir -p gpurun_out
timeout 100 python -m pytest tests/test_gpu_solver.py -x -q -k "two_iterations or time_steps_with_two" > gpurun_out/b_pytest.log 2>&1
echo "pytest rc=$?" >> gpurun_out/b_pytest.log
timeout 150 python tools/sweep_pt.py --grids 255x153x153 --modes FASTEST --zchunks 16,19,26 --tb2ty 8,16,32 \
    --sets "tb2_slim=0;tb2_slim=1,tb2_np=0,tb2_pf=0;tb2_slim=1,tb2_np=0,tb2_pf=1;tb2_slim=1,tb2_np=1,tb2_pf=0;tb2_slim=1,tb2_np=1,tb2_pf=1;tb2_slim=1,tb2_np=1,tb2_pf=2" \
    > gpurun_out/b_sweep_B.jsonl 2> gpurun_out/b_sweep_B.err
timeout 60 python tools/sweep_pt.py --grids 255x153x153 --modes FAST,PARITY --zchunks 16 --tb2ty 16 \
    --sets "tb2_slim=0;tb2_slim=1,tb2_np=1,tb2_pf=0;tb2_slim=1,tb2_np=1,tb2_pf=2" > gpurun_out/b_sweep_B_fast.jsonl 2>> gpurun_out/b_sweep_B.err
timeout 150 python tools/sweep_pt.py --grids 511x511x511 --modes FASTEST --zchunks 16,32,64 --tb2ty 16,32 --iters 60 --reps 2 \
    --sets "tb2_slim=0;tb2_slim=1,tb2_np=0,tb2_pf=0;tb2_slim=1,tb2_np=1,tb2_pf=0;tb2_slim=1,tb2_np=1,tb2_pf=2" > gpurun_out/b_sweep_511.jsonl 2> gpurun_out/b_sweep_511.err
timeout 90 ncu --set full --clock-control none --import-source on -k regex:pt_tb2s -c 2 -f -o gpurun_out/pt_tb2s_np1_pf0_B_fastest \
    python tools/profile_pt.py 255x153x153 FASTEST 0 1 tb2_np=1,tb2_pf=0 > gpurun_out/b_ncu1.log 2>&1
timeout 90 ncu --set full --clock-control none --import-source on -k regex:pt_tb2s -c 2 -f -o gpurun_out/pt_tb2s_np1_pf2_B_fastest \
    python tools/profile_pt.py 255x153x153 FASTEST 0 1 tb2_np=1,tb2_pf=2 > gpurun_out/b_ncu2.log 2>&1
tail -2 gpurun_out/b_pytest.log
for f in gpurun_out/b_sweep_B.jsonl gpurun_out/b_sweep_511.jsonl; do python -c "
import sys, json
rows=[json.loads(l) for l in open('$f') if l.strip()]
rows.sort(key=lambda r: r['us_per_iter'])
for r in rows[:6]: print(r['tb2_ty'], r['zchunk'], r['opts'], r['us_per_iter'], r['T_eff_GBs'])
"; done
