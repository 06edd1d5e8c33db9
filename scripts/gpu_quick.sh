#!/bin/bash
mkdir -p gpurun_out
for DT in tc fp32; do
timeout 600 python bench.py --workload R-lit --frames 20 --steps 3 --warmup 3 --precision $DT --no-cpu-baseline --no-e2e > gpurun_out/quick_$DT.log 2>&1
python - <<PY
import json
try:
    l=[x for x in open('gpurun_out/quick_$DT.log') if x.startswith('{')][-1]; d=json.loads(l)
    print('$DT', 'fps', round(d['value'],1), 'ms/frame', round(d['ms_per_step']/20,3), 'launches', d['gpu_launches'])
except Exception as e:
    print('$DT failed', e); print(open('gpurun_out/quick_$DT.log').read()[-1500:])
PY
done
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_tc.csv python bench.py --workload R-lit --frames 3 --steps 1 --warmup 1 --precision tc --no-cpu-baseline --no-e2e > gpurun_out/ncu_launch_tc.log 2>&1
wc -l gpurun_out/launches_tc.csv
