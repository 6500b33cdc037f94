#!/bin/bash
# the default bench line with z-bands, and the DRAM bytes of every ptv_kernel launch of 6 passes (small outputs only)
rm -rf gpurun_out/*; mkdir -p gpurun_out/r2c23 && cd "$(dirname "$0")/../.." || exit 1
O=gpurun_out/r2c23
timeout 500 python bench.py > $O/bench_B.json 2> $O/bench_B.err; echo "bench B rc=$?"; cut -c1-260 $O/bench_B.json; tail -2 $O/bench_B.err
timeout 200 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none -k regex:ptv_kernel -c 48 --csv --log-file $O/ptv_bands_dram.csv python tools/profile_pt.py 255x153x153 FAST 0 1 ptv_k=2 12 > $O/ncu_bands.log 2>&1; echo "ncu rc=$?"; tail -3 $O/ptv_bands_dram.csv | cut -c1-200
du -sh gpurun_out
echo "elapsed ${SECONDS}s"
