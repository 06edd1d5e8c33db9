#!/bin/bash
mkdir -p gpurun_out
O=gpurun_out
timeout 600 python -m pytest tests/test_gpu_half.py tests/test_gpu_ops.py -m gpu -q -p no:cacheprovider -rP 2>&1 | grep -E "half\]|passed|failed|Error|error" | head -30
for prec in tc half; do
timeout 600 python bench.py --precision $prec --steps 3 --warmup 3 --no-cpu-baseline --no-e2e --no-extras > $O/r2h_bench_$prec.log 2>&1
python -c "
import json
try:
    d=json.loads(open('$O/r2h_bench_$prec.log').read().strip().splitlines()[-1]); print('$prec', round(d['value'],1), 'fps', d['dtype'])
except Exception as e: print('$prec failed', e); print(open('$O/r2h_bench_$prec.log').read()[-1500:])
"
done
