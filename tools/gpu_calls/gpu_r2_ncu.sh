#!/bin/bash
# ncu --set full captures of ptv_kernel at config B for the option sets given as arguments (name=value,name=value ...)
mkdir -p gpurun_out/r2ncu && cd "$(dirname "$0")/../.." || exit 1
O=gpurun_out/r2ncu
MODE=${MODE:-FASTEST}
GRID=${GRID:-255x153x153}
i=0
for opts in "$@"; do
  i=$((i+1))
  python tools/profile_pt.py $GRID $MODE 0 1 "$opts" > $O/plain_$i.log 2>&1 &&
  timeout 200 ncu --set full --clock-control none --import-source on -k regex:ptv_kernel -s 1 -c 2 -f -o $O/ptv_${TAG:-x}_$i \
      python tools/profile_pt.py $GRID $MODE 0 1 "$opts" > $O/ncu_$i.log 2>&1
  echo "capture $i ($opts) rc=$?"; tail -2 $O/ncu_$i.log
done
echo "elapsed ${SECONDS}s"
