#!/bin/bash
mkdir -p gpurun_out
TAG=${TAG:-r2l}
(timeout 900 python -m pytest tests/test_gpu_zz_training.py -q --tb=short -s -rxX  > gpurun_out/${TAG}_train_tests_gpu.log 2>&1; echo "pytest exit $?" >> gpurun_out/${TAG}_train_tests_gpu.log)
grep -E "relative L2|passed|failed|Error|exit" gpurun_out/${TAG}_train_tests_gpu.log | tail

timeout 120 python scripts/bench_train.py --shape v7 --steps 5 --warmup 4 --graphs > gpurun_out/${TAG}_train_bench_v7.json 2> gpurun_out/${TAG}_train_err.txt
python -c "
import json; d=json.loads(open('gpurun_out/${TAG}_train_bench_v7.json').read().strip().splitlines()[-1]); print('v7', round(d['value'],1), 'fps', round(d['ms_per_step'],2), 'ms')"
