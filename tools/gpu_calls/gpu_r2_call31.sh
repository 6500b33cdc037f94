#!/bin/bash
# bench.py's new warm-up check against the script's text, and smoke(), on hardware
mkdir -p gpurun_out/r2c31 && cd "$(dirname "$0")/../.." || exit 1
timeout 60 python bench.py --steps 2 --warmup 2 --no-extras --no-cpu-baseline > gpurun_out/r2c31/bench_B_short.json 2> gpurun_out/r2c31/bench_B_short.err; echo "bench rc=$?"
python - <<'P'
import json
l=json.loads(open("gpurun_out/r2c31/bench_B_short.json").read().strip().splitlines()[-1])
print(l["value"], l["parity_check"]["reference_text"], l["parity_check"]["bit_identical"], l["pt_iters_per_step"])
P
timeout 40 python __graft_entry__.py --smoke 2>&1 | tail -3 | tee gpurun_out/r2c31/smoke.log
echo "elapsed ${SECONDS}s"
