#!/bin/bash
mkdir -p gpurun_out/r2c27 && cd "$(dirname "$0")/../.." || exit 1
timeout 100 python -m pytest tests/test_gpu_solver.py -m gpu -x -q -k "bands" > gpurun_out/r2c27/pytest_bands.log 2>&1; echo "pytest rc=$?"; tail -2 gpurun_out/r2c27/pytest_bands.log
echo "elapsed ${SECONDS}s"
