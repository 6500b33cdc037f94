#!/bin/bash
# Round 2, GPU call 1 (one B200): GPU test suite incl. the config-B whole-step parity test, the bench line with
# its extra legs (variant M at N=1, config E), configs C and D through bench.py, and the four round-1 candidates.
mkdir -p gpurun_out/r2c1 && cd "$(dirname "$0")/.." || exit 1
O=gpurun_out/r2c1
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > $O/gpu.txt 2>&1
timeout 600 python -m pytest tests -m gpu -x -q > $O/pytest.log 2>&1; echo "pytest rc=$?"; tail -3 $O/pytest.log
timeout 400 python bench.py > $O/bench_B.json 2> $O/bench_B.err; echo "bench B rc=$?"; cut -c1-300 $O/bench_B.json
timeout 200 python bench.py --workload C --pt-only 1000 --steps 5 --no-cpu-baseline > $O/bench_C.json 2> $O/bench_C.err; echo "bench C rc=$?"; cut -c1-400 $O/bench_C.json
timeout 200 python bench.py --workload C --pt-only 1000 --steps 5 --no-cpu-baseline --mode FASTEST > $O/bench_C_fastest.json 2> $O/bench_C_fastest.err; cut -c1-200 $O/bench_C_fastest.json
timeout 400 python bench.py --workload D --steps 5 --warmup 1 --no-parity-check --no-e2e --no-cpu-baseline --no-extras > $O/bench_D.json 2> $O/bench_D.err; echo "bench D rc=$?"; cut -c1-400 $O/bench_D.json
timeout 200 python tools/sweep_pt.py --grids 255x153x153 --modes FASTEST,FAST --zchunks 0 --tb2ty 0 --iters 304 \
   --sets "tb2_dual=0;tb2_dual=2;tb2_dual=0,tb2_pairbar=1;tb2_pairbar=0,pt_bands=4;pt_bands=2;pt_bands=0" > $O/cand_B.jsonl 2> $O/cand_B.err; echo "cand B rc=$?"; cat $O/cand_B.jsonl
timeout 200 python tools/sweep_pt.py --grids 511x511x511 --modes FASTEST --zchunks 0 --tb2ty 0 --iters 100 \
   --sets "tb2_dual=0;tb2_dual=2;tb2_dual=0,tb2_pairbar=1;tb2_pairbar=0,pt_bands=4;pt_bands=0" > $O/cand_511.jsonl 2> $O/cand_511.err; echo "cand 511 rc=$?"; cat $O/cand_511.jsonl
echo "elapsed ${SECONDS}s"
