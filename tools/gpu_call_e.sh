#!/bin/bash
# Round-1 GPU call E: compile-time-stride instantiations of the slim kernel: parity, timing against the generic one, ncu.
mkdir -p gpurun_out
timeout 200 python -m pytest tests/test_gpu_solver.py -x -q -k "compile_time_stride or two_iterations or spot_check" > gpurun_out/e_pytest.log 2>&1
echo "pytest rc=$?" >> gpurun_out/e_pytest.log
timeout 100 python tools/sweep_pt.py --grids 255x153x153 --modes FASTEST,FAST --zchunks 12,16 --tb2ty 8 \
    --sets "tb2_spec=0;tb2_spec=1" > gpurun_out/e_sweep_B.jsonl 2> gpurun_out/e_sweep_B.err
timeout 120 python tools/sweep_pt.py --grids 511x511x511 --modes FASTEST --zchunks 32,64 --tb2ty 16 --iters 60 --reps 2 \
    --sets "tb2_spec=0;tb2_spec=1" > gpurun_out/e_sweep_511.jsonl 2> gpurun_out/e_sweep_511.err
timeout 90 ncu --set full --clock-control none --import-source on -k regex:pt_tb2s -c 2 -f -o gpurun_out/pt_tb2s_spec_B_fastest \
    python tools/profile_pt.py 255x153x153 FASTEST 0 1 > gpurun_out/e_ncu1.log 2>&1
tail -3 gpurun_out/e_pytest.log
for f in gpurun_out/e_sweep_B.jsonl gpurun_out/e_sweep_511.jsonl; do python -c "
import sys, json
rows=[json.loads(l) for l in open('$f') if l.strip()]
for r in rows: print(r['mode'], r['tb2_ty'], r['zchunk'], r['opts'], r['us_per_iter'], r['T_eff_GBs'])
"; done
