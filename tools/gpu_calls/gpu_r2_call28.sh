#!/bin/bash
# the library on hardware against the reference-text fixtures (8 whole runs through ns3d_step / groups / level 1) + smoke
mkdir -p gpurun_out/r2c28 && cd "$(dirname "$0")/../.." || exit 1
timeout 80 python -m pytest tests/test_gpu_driver.py -m gpu -x -q -k "reference_text" > gpurun_out/r2c28/pytest_text.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/r2c28/pytest_text.log
echo "elapsed ${SECONDS}s"
