#!/bin/bash
mkdir -p gpurun_out
TAG=${TAG:-r2v}
B="python bench.py --steps 4 --warmup 3 --no-cpu-baseline --no-extras --no-e2e"
for v in "" "CRFP_TC3_NOFAST16=1"; do
  env $v timeout 600 $B > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err
  python -c "
import json; d=json.loads(open('gpurun_out/${TAG}_bench.json').read().strip().splitlines()[-1]); print('[$v] value', round(d['value'],1), 'conv frac', round(d['roofline']['frac'],3), 'avg launch us', round(d['roofline']['avg_launch_ms']*1e3,1))"
done
python scripts/tc3_ws_trace.py 2>/dev/null | grep -E "^==|rows " | head -8
