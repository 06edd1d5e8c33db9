#!/bin/bash
# tests + tc-precision quick benches at R-lit and R-nat (+ optional env for A/B)
mkdir -p gpurun_out
timeout 500 python -m pytest tests -m gpu -q 2>&1 | tail -4
for WL in R-lit R-nat; do
timeout 300 python bench.py --workload $WL --frames 20 --steps 3 --warmup 3 --no-cpu-baseline --no-e2e 2>&1 | python -c "
import sys,json
ls=[l for l in sys.stdin if l.startswith(chr(123))]
if not ls: print('$WL FAILED'); sys.exit(0)
d=json.loads(ls[-1]); print('$WL', round(d['value'],1), 'fps', round(d['ms_per_step']/20,3), 'ms/frame | conv frac', round(d['roofline']['frac'],3), 'align frac', round(d['roofline']['align_kernel']['frac'],3), 'align ms', round(d['roofline']['align_kernel']['avg_launch_ms'],4))"
done
