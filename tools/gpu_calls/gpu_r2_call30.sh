#!/bin/bash
# the library at the benchmark configuration against the record of the shipped script's text (2 time steps, PARITY and FAST)
mkdir -p gpurun_out/r2c30 && cd "$(dirname "$0")/../.." || exit 1
timeout 80 python -m pytest tests/test_gpu_solver.py -m gpu -q -k "shipped_scripts_text" > gpurun_out/r2c30/pytest_config_B_text.log 2>&1; echo "pytest rc=$?"; tail -6 gpurun_out/r2c30/pytest_config_B_text.log
echo "elapsed ${SECONDS}s"
