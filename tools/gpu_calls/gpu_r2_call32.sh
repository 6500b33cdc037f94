#!/bin/bash
# the library at config B, variant M, against the record of the multi-GPU script's text (3 time steps, PARITY)
mkdir -p gpurun_out/r2c32 && cd "$(dirname "$0")/../.." || exit 1
timeout 14 python -m pytest tests/test_gpu_solver.py -m gpu -q -k "variant_M_vs_the_multi" -p no:cacheprovider > gpurun_out/r2c32/pytest_config_B_M_text.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/r2c32/pytest_config_B_M_text.log
