#!/bin/bash
# Round 2, multi-GPU call (N = $1 GPUs of one box): the slab parity tests on hardware, then the bench line at N.
N=${1:-2}
mkdir -p gpurun_out/r2mg$N && cd "$(dirname "$0")/.." || exit 1
O=gpurun_out/r2mg$N
nvidia-smi --query-gpu=index,name,clocks.sm --format=csv > $O/gpu.txt 2>&1
timeout 900 python -m pytest tests/test_gpu_multi.py -m gpu -q -rs > $O/pytest_multi.log 2>&1; echo "pytest multi rc=$?"; tail -15 $O/pytest_multi.log
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533"
timeout 600 $TR bench.py --gpus $N > $O/bench_n$N.json 2> $O/bench_n$N.err; echo "bench rc=$?"; cut -c1-300 $O/bench_n$N.json
if [ -n "$2" ]; then
  timeout 400 $TR bench.py --gpus $N --no-extras --no-e2e --opt tb2_slim_faces=1 > $O/bench_n${N}_slimfaces.json 2> $O/bench_n${N}_slimfaces.err; echo "slim rc=$?"; cut -c1-300 $O/bench_n${N}_slimfaces.json
fi
echo "elapsed ${SECONDS}s"
