#!/bin/bash
# multi-GPU evidence: N=${N} -- the slab parity tests, then the bench line (incl. the config-E leg)
N=${N:-2}
mkdir -p gpurun_out/r2mg${N}b && cd "$(dirname "$0")/../.." || exit 1
O=gpurun_out/r2mg${N}b
nvidia-smi --query-gpu=index,name --format=csv > $O/gpu.txt
if [ -z "$SKIP_TESTS" ]; then
  timeout 900 python -m pytest tests/test_gpu_multi.py -m gpu -q -x > $O/pytest_multi.log 2>&1; echo "pytest multi rc=$?"; tail -3 $O/pytest_multi.log
fi
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus $N --steps 5 --warmup 3 > $O/bench_n$N.json 2> $O/bench_n$N.err; echo "bench rc=$?"; tail -1 $O/bench_n$N.json | cut -c1-400
echo "elapsed ${SECONDS}s"
