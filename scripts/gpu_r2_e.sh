#!/bin/bash
mkdir -p gpurun_out
O=gpurun_out
timeout 300 python -m pytest tests/test_gpu_umma.py -m gpu -q -x -p no:cacheprovider -rP 2>&1 | tail -15
timeout 300 python -m pytest tests/test_gpu_tc.py -m gpu -q -x -p no:cacheprovider -k "fused" -rP > $O/r2e_fused_tests.log 2>&1
echo "fused tests rc $?"; tail -30 $O/r2e_fused_tests.log
