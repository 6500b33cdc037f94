#!/bin/bash
# Round-1 GPU call A: parity subset of the slim two-iteration kernel, variant sweep, one ncu capture, short bench.
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.max.sm,power.limit --format=csv,noheader > gpurun_out/a_gpu.txt 2>&1
timeout 150 python -m pytest tests/test_gpu_solver.py -x -q -k "two_iterations or fused_iteration or time_steps_with_two" > gpurun_out/a_pytest.log 2>&1
echo "pytest rc=$?" >> gpurun_out/a_pytest.log
timeout 120 python tools/sweep_pt.py --grids 255x153x153 --modes FASTEST --zchunks 16,19,26,38 --tb2ty 8,16,32 \
    --sets "tb2_slim=0,tb2_pf=0;tb2_slim=1,tb2_pf=0;tb2_slim=1,tb2_pf=1" > gpurun_out/a_sweep_B.jsonl 2> gpurun_out/a_sweep_B.err
timeout 60 python tools/sweep_pt.py --grids 255x153x153 --modes FAST --zchunks 16,19 --tb2ty 16 \
    --sets "tb2_slim=0,tb2_pf=0;tb2_slim=1,tb2_pf=0;tb2_slim=1,tb2_pf=1" > gpurun_out/a_sweep_B_fast.jsonl 2>> gpurun_out/a_sweep_B.err
timeout 120 python tools/sweep_pt.py --grids 511x511x511 --modes FASTEST --zchunks 16,32,64 --tb2ty 16,32 --iters 60 --reps 2 \
    --sets "tb2_slim=0,tb2_pf=0;tb2_slim=1,tb2_pf=0;tb2_slim=1,tb2_pf=1" > gpurun_out/a_sweep_511.jsonl 2> gpurun_out/a_sweep_511.err
timeout 120 ncu --set full --clock-control none --import-source on -k regex:pt_tb2s -c 2 -f -o gpurun_out/pt_tb2s_B_fastest \
    python tools/profile_pt.py 255x153x153 FASTEST 0 1 > gpurun_out/a_ncu.log 2>&1
timeout 150 python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/a_bench.json 2> gpurun_out/a_bench.err
tail -3 gpurun_out/a_pytest.log
cat gpurun_out/a_sweep_B.jsonl | python -c "
import sys, json
rows=[json.loads(l) for l in sys.stdin if l.strip()]
rows.sort(key=lambda r: r['us_per_iter'])
for r in rows[:8]: print(r)
"
cut -c1-400 gpurun_out/a_bench.json
