#!/bin/bash
# diagnose the launch failure of ptv_flow_kernel<.,256,2> (k2_lb0, 63x38x38): sanitizer on the failing case, then the whole GPU suite
mkdir -p gpurun_out/r2c12 && cd "$(dirname "$0")/../.." || exit 1
O=gpurun_out/r2c12
timeout 300 python -m pytest tests/test_gpu_solver.py -m gpu -q -k "test_ptv_kernel_bit_exact and k2_lb0" > $O/pytest_k2lb0.log 2>&1; echo "k2_lb0 rc=$?"; tail -4 $O/pytest_k2lb0.log
timeout 500 compute-sanitizer --tool memcheck --error-exitcode 7 python -m pytest tests/test_gpu_solver.py -m gpu -x -q -k "test_ptv_kernel_bit_exact and k2_lb0 and grid5" > $O/memcheck.log 2>&1; echo "memcheck rc=$?"; grep -v "^\." $O/memcheck.log | head -60
timeout 900 python -m pytest tests -m gpu -q > $O/pytest_all.log 2>&1; echo "pytest all rc=$?"; tail -8 $O/pytest_all.log
echo "elapsed ${SECONDS}s"
