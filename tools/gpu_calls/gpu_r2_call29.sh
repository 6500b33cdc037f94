#!/bin/bash
# the Julia shim's text + the re-pointed run scripts, interpreted, every ccall into libns3d.so on the B200
mkdir -p gpurun_out/r2c29 && cd "$(dirname "$0")/../.." || exit 1
timeout 90 python -m pytest tests/test_julia_shim_exec.py -m gpu -q -rs > gpurun_out/r2c29/pytest_shim.log 2>&1; echo "pytest rc=$?"; tail -6 gpurun_out/r2c29/pytest_shim.log
echo "elapsed ${SECONDS}s"
