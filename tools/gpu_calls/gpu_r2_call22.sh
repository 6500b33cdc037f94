#!/bin/bash
# final single-GPU evidence with z-bands: GPU suite, the default bench line, ncu of every ptv_kernel launch of 6 passes
mkdir -p gpurun_out/r2c22 gpurun_out/r2ncu && cd "$(dirname "$0")/../.." || exit 1
O=gpurun_out/r2c22
nvidia-smi --query-gpu=name,clocks.max.sm,power.limit --format=csv > $O/gpu.txt
timeout 600 python -m pytest tests -m gpu -q -x > $O/pytest_all.log 2>&1; echo "pytest all rc=$?"; tail -3 $O/pytest_all.log
timeout 500 python bench.py > $O/bench_B.json 2> $O/bench_B.err; echo "bench B rc=$?"; cut -c1-260 $O/bench_B.json; tail -2 $O/bench_B.err
timeout 300 ncu --set full --clock-control none --import-source on -k regex:ptv_kernel -c 48 -f -o gpurun_out/r2ncu/ptv_bands_B_FAST python tools/profile_pt.py 255x153x153 FAST 0 1 ptv_k=2 12 > $O/ncu_bands.log 2>&1; echo "ncu rc=$?"; tail -1 $O/ncu_bands.log
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 1500 --csv --log-file $O/launches.csv python bench.py --steps 1 --warmup 1 --no-extras --no-cpu-baseline --no-e2e --no-parity-check --fixed-iters 304 > $O/ncu_launch.log 2>&1; echo "launch list rc=$?"
python tools/launch_list_summary.py $O/launches.csv $O/launch_summary.csv "bench.py --steps 1 --warmup 1 --fixed-iters 304 (workload B, FAST, z-bands)" | head -20
rm -f $O/launches.csv
echo "elapsed ${SECONDS}s"
