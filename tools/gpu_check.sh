#!/bin/bash
# One GPU call that produces what a round needs: the bench line, the launch list, ncu captures of the default
# kernel, then the GPU test suite (BUDGET = seconds of run time the call may use; the tail is optional).
mkdir -p gpurun_out
BUDGET=${BUDGET:-175}
timeout 150 python bench.py > gpurun_out/g_bench_n1.json 2> gpurun_out/g_bench_n1.err
echo "bench rc=$?"
cut -c1-600 gpurun_out/g_bench_n1.json
timeout 60 ncu --set full --clock-control none --import-source on -k regex:pt_tb2s -c 2 -f -o gpurun_out/pt_tb2s_default_B_fastest \
    python tools/profile_pt.py 255x153x153 FASTEST 0 1 > gpurun_out/g_ncu_B.log 2>&1
timeout 90 ncu --metrics gpu__time_duration.sum --clock-control none -c 2500 --csv --log-file gpurun_out/g_launches.csv \
    python bench.py --steps 1 --warmup 3 --no-e2e --no-cpu-baseline --no-parity-check > gpurun_out/g_ncu_launch.log 2>&1
python tools/launch_list_summary.py gpurun_out/g_launches.csv gpurun_out/g_launch_summary.csv \
    "ncu --metrics gpu__time_duration.sum --clock-control none -c 2500 python bench.py --steps 1 --warmup 3 --no-e2e --no-cpu-baseline --no-parity-check" | head -14
timeout $((BUDGET - SECONDS - 3)) python -m pytest tests -m gpu -x -q > gpurun_out/g_pytest.log 2>&1
echo "pytest rc=$?"
tail -4 gpurun_out/g_pytest.log
echo "elapsed ${SECONDS}s"
# optional tail, only while the call's budget lasts (BUDGET = seconds of run time this call may use)
if [ $((BUDGET - SECONDS)) -gt 35 ]; then
  timeout $((BUDGET - SECONDS - 5)) python bench.py --workload C --pt-only 600 --steps 2 --no-cpu-baseline > gpurun_out/g_bench_C.json 2> gpurun_out/g_bench_C.err
  cut -c1-400 gpurun_out/g_bench_C.json
fi
if [ $((BUDGET - SECONDS)) -gt 40 ]; then
  timeout $((BUDGET - SECONDS - 5)) ncu --set full --clock-control none --import-source on -k regex:pt_tb2s -c 2 -f -o gpurun_out/pt_tb2s_default_511_fastest \
      python tools/profile_pt.py 511x511x511 FASTEST 0 1 > gpurun_out/g_ncu_511.log 2>&1
fi
echo "elapsed ${SECONDS}s"
