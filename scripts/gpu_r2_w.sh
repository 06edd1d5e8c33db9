#!/bin/bash
mkdir -p gpurun_out
(timeout 600 python -m pytest tests/test_gpu_half.py tests/test_gpu_tc.py -q -x --tb=short 2>&1 | tail -2)
for p in half tc; do
timeout 600 python bench.py --precision $p --steps 4 --warmup 3 --no-cpu-baseline --no-extras --no-e2e > gpurun_out/r2w_bench_$p.json 2> gpurun_out/r2w_bench.err
python -c "
import json; d=json.loads(open('gpurun_out/r2w_bench_$p.json').read().strip().splitlines()[-1]); print('$p value', round(d['value'],1))"
done
