#!/bin/bash
# A/B of env switches on the quick R-lit bench: usage  AB="CRFP_DCN_V1=1 CRFP_NO_PDL=1" bash scripts/gpu_ab.sh
run() {
  env $1 timeout 300 python bench.py --workload ${WL:-R-lit} --frames 20 --steps 3 --warmup 3 --no-cpu-baseline --no-e2e 2>&1 | python -c "
import sys,json
ls=[l for l in sys.stdin if l.startswith(chr(123))]
if not ls: print('$1 FAILED'); sys.exit(0)
d=json.loads(ls[-1]); print('$1', round(d['value'],1), 'fps', round(d['ms_per_step']/20,3), 'ms/frame | conv frac', round(d['roofline']['frac'],3), 'align frac', round(d['roofline']['align_kernel']['frac'],3), 'align ms', round(d['roofline']['align_kernel']['avg_launch_ms'],4), d['clocks'])"
}
run X=0
for v in $AB; do run $v; done
run X=0
