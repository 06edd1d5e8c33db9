#!/bin/bash
mkdir -p gpurun_out
TAG=${TAG:-r2p}
(timeout 900 python -m pytest tests/test_gpu_ops.py tests/test_gpu_tc.py tests/test_gpu_model.py tests/test_gpu_long.py -q --tb=short -x > gpurun_out/${TAG}_tests.log 2>&1; echo "pytest exit $?" >> gpurun_out/${TAG}_tests.log)
tail -3 gpurun_out/${TAG}_tests.log
B="python bench.py --steps 4 --warmup 3 --no-cpu-baseline --no-extras --no-e2e"
for v in "" "CRFP_PDL=all"; do
  env $v timeout 600 $B > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err
  python -c "
import json; d=json.loads(open('gpurun_out/${TAG}_bench.json').read().strip().splitlines()[-1]); print('[$v] value', round(d['value'],1))"
done
CRFP_NO_GRAPHS=1 CRFP_PDL=none timeout 600 python scripts/kernel_times.py --frames 20 --steps 4 2>/dev/null | tail -18 | head -9
