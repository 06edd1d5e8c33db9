#!/bin/bash
mkdir -p gpurun_out
export CRFP_TC3_NOCOAL=1 CRFP_TC3_NOCONST=1
for f in 0 4; do
  echo "#### CRFP_TC3_DBG=$f"
  CRFP_TC3_DBG=$f timeout 300 python scripts/tc3_ws_trace.py 2>&1 | grep -E "^==|rows |epilogue warp"
done
