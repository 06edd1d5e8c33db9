#!/bin/bash
# Round 2, GPU call A: full GPU test suite, headline bench (new fields), A/B of the new switches, in-pipeline kernel
# times, conv pipeline trace, compute-sanitizer on the smoke shapes.  Everything under gpurun_out/.
mkdir -p gpurun_out
TAG=${TAG:-r2a}
O=gpurun_out
nvidia-smi --query-gpu=name,clocks.max.sm,memory.total --format=csv > $O/${TAG}_gpu.txt 2>&1
lscpu | head -20 >> $O/${TAG}_gpu.txt; numactl -H >> $O/${TAG}_gpu.txt 2>&1
timeout 1500 python -m pytest tests -m gpu -q -rP --durations=12 -p no:cacheprovider > $O/${TAG}_tests.log 2>&1
echo "tests rc $?" >> $O/${TAG}_tests.log
tail -5 $O/${TAG}_tests.log
timeout 900 python bench.py --steps 3 --warmup 3 > $O/${TAG}_bench.log 2> $O/${TAG}_bench.err
echo "bench rc $?"; tail -c 600 $O/${TAG}_bench.err
python - << 'PY'
import json
try:
    d = json.loads(open("gpurun_out/%s_bench.log" % __import__("os").environ.get("TAG", "r2a")).read().strip().splitlines()[-1])
    print("value", round(d["value"], 1), "e2e", round(d["e2e"]["value"], 1), "wall", round(d["e2e"]["wall_value"], 1),
          "roof", round(d["roofline"]["frac"], 3), "align", round(d["roofline"]["align_kernel"]["frac"], 3),
          "align_ms", round(d["roofline"]["align_kernel"]["avg_launch_ms"], 4), "conv_ms", round(d["roofline"]["avg_launch_ms"], 4),
          "cpu", d["cpu_baseline"]["value"], "stock", d["gpu_stock_baseline"], "extra", [(e["workload"][:5], round(e["value"], 1)) for e in d["extra"]])
except Exception as e:
    print("bench parse failed", e)
PY
ab() {  # env assignments..., label
  label=$1; shift
  env "$@" timeout 600 python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-e2e --no-extras > $O/${TAG}_ab_$label.log 2>&1
  python -c "
import json,sys
try:
    d=json.loads(open('$O/${TAG}_ab_$label.log').read().strip().splitlines()[-1])
    print('$label', round(d['value'],1), 'fps; align_ms', round(d['roofline']['align_kernel']['avg_launch_ms'],4), 'conv_ms', round(d['roofline']['avg_launch_ms'],4))
except Exception as e: print('$label failed', e)
"
}
ab default X=1
ab noaux CRFP_AUX=0
ab headepi CRFP_HEAD_EPI=1
ab noaux_headepi CRFP_AUX=0 CRFP_HEAD_EPI=1
timeout 600 python bench.py --clips 2 --frames 40 --steps 3 --warmup 3 --no-cpu-baseline --no-e2e --no-extras > $O/${TAG}_clips2.log 2>&1
python -c "
import json
try:
    d=json.loads(open('$O/${TAG}_clips2.log').read().strip().splitlines()[-1]); print('clips=2 frames=40:', round(d['value'],1), 'fps')
except Exception as e: print('clips2 failed', e)
"
timeout 600 python scripts/kernel_times.py --frames 20 --steps 6 > $O/${TAG}_kernel_times.txt 2>&1
tail -22 $O/${TAG}_kernel_times.txt
timeout 300 python scripts/tc3_ws_trace.py > $O/${TAG}_tc3_trace.txt 2>&1
HEADS=1 timeout 300 python scripts/tc3_ws_trace.py >> $O/${TAG}_tc3_trace.txt 2>&1
grep -E "^==|rows " $O/${TAG}_tc3_trace.txt
timeout 900 compute-sanitizer --tool memcheck --error-exitcode 7 python -c "import __graft_entry__ as g; g.smoke()" > $O/${TAG}_sanitizer_memcheck.log 2>&1
echo "memcheck rc $?"; tail -4 $O/${TAG}_sanitizer_memcheck.log
ls -la $O | head -40; du -sh $O
