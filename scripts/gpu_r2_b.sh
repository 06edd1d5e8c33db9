#!/bin/bash
# quick check of a kernel change: tensor-core op tests + model parity, A/B bench lines, conv trace
mkdir -p gpurun_out
TAG=${TAG:-r2b}
O=gpurun_out
timeout 900 python -m pytest tests/test_gpu_tc.py tests/test_gpu_model.py tests/test_gpu_long.py -m gpu -q -x -p no:cacheprovider > $O/${TAG}_tests.log 2>&1
echo "tests rc $?"; tail -3 $O/${TAG}_tests.log
ab() {
  label=$1; shift
  env "$@" timeout 600 python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-e2e --no-extras > $O/${TAG}_ab_$label.log 2>&1
  python -c "
import json
try:
    d=json.loads(open('$O/${TAG}_ab_$label.log').read().strip().splitlines()[-1])
    print('$label', round(d['value'],1), 'fps; align_ms', round(d['roofline']['align_kernel']['avg_launch_ms'],4), 'conv_ms', round(d['roofline']['avg_launch_ms'],4), 'conv frac', round(d['roofline']['frac'],3))
except Exception as e: print('$label failed', e)
"
}
ab default X=1
for v in $ABS; do ab $(echo $v | tr '=,' '__') $(echo $v | tr ',' ' '); done
timeout 300 python scripts/tc3_ws_trace.py > $O/${TAG}_tc3_trace.txt 2>&1
grep -E "^==|rows " $O/${TAG}_tc3_trace.txt
