#!/bin/bash
mkdir -p gpurun_out/r2c25 && cd "$(dirname "$0")/../.." || exit 1
O=gpurun_out/r2c25
run() { timeout 200 python bench.py --steps 4 --warmup 3 --no-e2e --no-extras --no-cpu-baseline "$@" 2>/dev/null | tail -1 | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print(sys.argv[1:], 'T_eff', round(d['value'],1), 'ms/step', round(d['ms_per_step'],2), 'us/pass', round(d['roofline']['us_per_launch'],2), 'share', round(d['roofline']['share_of_step'],3), 'bit-identical', d.get('parity_check',{}).get('bit_identical'))" "$@"; }
{
run --opt graph_pieces=4
run --opt graph_pieces=1 --no-parity-check
run --opt graph_pieces=8 --no-parity-check
run --opt graph_pieces=2 --no-parity-check
} 2>&1 | tee $O/pieces_bench.log
echo "elapsed ${SECONDS}s"
