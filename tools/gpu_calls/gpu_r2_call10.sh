#!/bin/bash
# ncu evidence for ptv_kernel: full captures at config B (FAST, FASTEST) and 511^3 (FAST), then the launch list of a bench run
mkdir -p gpurun_out/r2c10 && cd "$(dirname "$0")/../.." || exit 1
O=gpurun_out/r2c10
nvidia-smi --query-gpu=name,clocks.max.sm,power.limit --format=csv > $O/gpu.txt
MODE=FAST TAG=B_FAST bash tools/gpu_r2_ncu.sh ptv_k=2
MODE=FASTEST TAG=B_FASTEST bash tools/gpu_r2_ncu.sh ptv_k=2
MODE=FAST GRID=511x511x511 TAG=511_FAST bash tools/gpu_r2_ncu.sh ptv_k=2,zchunk=32
timeout 300 python bench.py --steps 2 --warmup 3 --no-extras --no-cpu-baseline > $O/bench_short.json 2> $O/bench_short.err; echo "bench rc=$?"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 6000 --csv --log-file $O/launches.csv python bench.py --steps 1 --warmup 1 --no-extras --no-cpu-baseline --no-e2e --no-parity-check > $O/ncu_launch.log 2>&1; echo "launch list rc=$?"
python tools/launch_list_summary.py $O/launches.csv $O/launch_summary.csv "bench.py --steps 1 --warmup 1 (workload B, FAST)"
rm -f $O/launches.csv
echo "elapsed ${SECONDS}s"
