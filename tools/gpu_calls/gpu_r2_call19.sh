#!/bin/bash
# final single-GPU evidence of the round: the GPU suite, the bench lines of configs B (default), C, D, E(N=1), ncu of ptv_kernel
mkdir -p gpurun_out/r2c19 gpurun_out/r2ncu && cd "$(dirname "$0")/../.." || exit 1
O=gpurun_out/r2c19
nvidia-smi --query-gpu=name,clocks.max.sm,power.limit --format=csv > $O/gpu.txt
timeout 600 python -m pytest tests -m gpu -q -x > $O/pytest_all.log 2>&1; echo "pytest all rc=$?"; tail -3 $O/pytest_all.log
timeout 500 python bench.py > $O/bench_B.json 2> $O/bench_B.err; echo "bench B rc=$?"; cut -c1-260 $O/bench_B.json; tail -2 $O/bench_B.err
timeout 200 python bench.py --workload C --pt-only 1000 --steps 5 --no-cpu-baseline > $O/bench_C.json 2> $O/bench_C.err; echo "bench C rc=$?"; cut -c1-500 $O/bench_C.json
timeout 400 python bench.py --workload D --steps 2 --warmup 1 --no-parity-check --no-e2e --no-cpu-baseline --no-extras > $O/bench_D.json 2> $O/bench_D.err; echo "bench D rc=$?"; cut -c1-260 $O/bench_D.json; tail -2 $O/bench_D.err
MODE=FAST TAG=final_B_FAST bash tools/gpu_r2_ncu.sh ptv_k=2
MODE=FAST GRID=511x511x511 TAG=final_511_FAST bash tools/gpu_r2_ncu.sh ptv_k=2
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:'predictor|corrector|advect|divV|pack|residual|bc_kernel|cylinder|fill' --csv --log-file $O/launches_once.csv python bench.py --steps 2 --warmup 1 --no-extras --no-cpu-baseline --no-e2e --no-parity-check > $O/ncu_launch.log 2>&1; echo "launch list rc=$?"
python tools/launch_list_summary.py $O/launches_once.csv $O/once_per_step_summary.csv "bench.py --steps 2 --warmup 1 (workload B, FAST): every kernel but ptv_kernel"
echo "elapsed ${SECONDS}s"
