#!/bin/bash
mkdir -p gpurun_out
O=gpurun_out
timeout 300 python -m pytest tests/test_gpu_tc.py -m gpu -q -x -p no:cacheprovider -k "fused" 2>&1 | tail -3
timeout 300 python scripts/fused_trace.py 2>&1 | tail -10
