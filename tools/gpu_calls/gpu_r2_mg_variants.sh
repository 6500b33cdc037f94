#!/bin/bash
# z-slabs: interface chunks as their own launch (default) against one launch per pass
N=${N:-2}
mkdir -p gpurun_out/r2mgv && cd "$(dirname "$0")/../.." || exit 1
O=gpurun_out/r2mgv
run() { timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus $N --steps 3 --warmup 3 --no-e2e --no-extras --no-cpu-baseline "$@" 2>/dev/null | tail -1 | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print(sys.argv[1:], 'per GPU', round(d['t_eff_per_gpu'],1), 'ms/step', round(d['ms_per_step'],2), 'us/pass', round(d['roofline']['us_per_launch'],2), 'parity', d['parity_check']['bit_identical'])" "$@"; }
{
run --opt p2p_split=1
run --opt p2p_split=0
run --opt p2p_split=0 --zchunk 8
run --workload E --fixed-iters 1020 --steps 2 --warmup 1 --no-parity-check --opt p2p_split=1
run --workload E --fixed-iters 1020 --steps 2 --warmup 1 --no-parity-check --opt p2p_split=0
} 2>&1 | tee $O/variants_n$N.log
echo "elapsed ${SECONDS}s"
