#!/bin/bash
# Round-2 artefact run (1 GPU): both bench arms, n=2/4 clip batches, streaming latency, ncu launch list + DRAM bytes of one
# steady-state frame, ncu --set full captures of the hot kernels (summarised on the box), compute-sanitizer (memcheck +
# racecheck at small shapes).  Everything under gpurun_out/r02_*.
mkdir -p gpurun_out
TAG=${TAG:-r02}
O=gpurun_out
if [ -z "$SKIP_BENCH" ]; then
timeout 900 python bench.py --impl reference --steps 2 --warmup 1 > $O/${TAG}_bench_reference.json 2> $O/${TAG}_bench_reference.err
timeout 1200 python bench.py --steps 5 --warmup 3 > $O/${TAG}_bench_R-lit.json 2> $O/${TAG}_bench.err
python - << 'PY'
import json, os
t = os.environ.get("TAG", "r02")
try:
    d = json.loads(open(f"gpurun_out/{t}_bench_R-lit.json").read().strip().splitlines()[-1])
    print("value", round(d["value"], 1), "e2e", round(d["e2e"]["value"], 1), "wall", round(d["e2e"]["wall_value"], 1), "f32", round(d["e2e_f32"]["value"], 1),
          "| conv frac", round(d["roofline"]["frac"], 3), "align", d["roofline"]["align_kernel"]["bound"], round(d["roofline"]["align_kernel"]["frac"], 3),
          round(d["roofline"]["align_kernel"]["avg_launch_ms"], 4), "ms | cpu", round(d["cpu_baseline"]["value"], 3), "stock", d["gpu_stock_baseline"].get("fp32_fps"), d["gpu_stock_baseline"].get("tf32_fps"))
    print("extra", [(e.get("workload", "")[:5], e.get("precision"), round(e.get("value", 0), 1)) for e in d["extra"]])
except Exception as e:
    print("bench parse failed", e)
PY
for n in 2 4; do
timeout 600 python bench.py --clips $n --frames 40 --steps 3 --warmup 3 --no-cpu-baseline --no-e2e --no-extras > $O/${TAG}_bench_clips$n.json 2>> $O/${TAG}_bench.err
python -c "
import json
try:
    d=json.loads(open('$O/${TAG}_bench_clips$n.json').read().strip().splitlines()[-1]); print('clips=$n x 40 frames:', round(d['value'],1), 'fps')
except Exception as e: print('clips$n failed', e)
"
done
timeout 300 python scripts/bench_stream.py > $O/${TAG}_stream_1080p.json 2>> $O/${TAG}_bench.err; cat $O/${TAG}_stream_1080p.json
fi
# ncu sees the individual launches: graphs off (a graph replay launches exactly these kernels)
export CRFP_NO_GRAPHS=1
BENCH="python bench.py --frames 3 --steps 1 --warmup 1 --no-cpu-baseline --no-e2e --no-extras"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file $O/${TAG}_launches_R-lit.csv $BENCH > $O/ncu_launch.log 2>&1
python scripts/launch_summary.py $O/${TAG}_launches_R-lit.csv > $O/${TAG}_launches_R-lit_summary.txt 2>&1; head -16 $O/${TAG}_launches_R-lit_summary.txt
timeout 900 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none --csv --log-file $O/${TAG}_dram_R-lit.csv $BENCH > $O/ncu_dram.log 2>&1
cap() {  # kernel regex, skip, count
  timeout 900 ncu --set full --clock-control none --import-source on -k regex:$1 -s $2 -c $3 -f -o $O/${TAG}_prof_$1 $BENCH > $O/ncu_$1.log 2>&1
  python scripts/ncu_summary.py $O/${TAG}_prof_$1.ncu-rep > $O/${TAG}_$1_ncu_full.txt 2>&1
  python scripts/ncu_hot.py $O/${TAG}_prof_$1.ncu-rep 30 >> $O/${TAG}_$1_ncu_full.txt 2>&1
  rm -f $O/${TAG}_prof_$1.ncu-rep
  tail -1 $O/ncu_$1.log
}
cap conv_tc3_ws_kernel 60 6
cap dcn_align_fused_kernel 3 2
cap conv_thin4t_kernel 30 4
cap dcn_hr_kernel 2 1
unset CRFP_NO_GRAPHS
timeout 900 compute-sanitizer --tool memcheck --error-exitcode 7 python -c "import __graft_entry__ as g; g.smoke()" > $O/${TAG}_sanitizer_memcheck.log 2>&1; echo "memcheck rc $?"; tail -3 $O/${TAG}_sanitizer_memcheck.log
timeout 900 compute-sanitizer --tool racecheck --error-exitcode 7 python -c "
import torch
from crfp_b200 import CRFP_DSV
from crfp_b200.synthetic import make_clip, make_state_dict
m = CRFP_DSV('cuda', mid_channels=32).eval(); m.load_state_dict(make_state_dict(seed=1), strict=True); m.cuda(); m.use_graphs = False
lrs, fvs, mks, _ = make_clip(seed=2, n=1, t=2, h=16, w=24, fv_size=48)
out = m(lrs.cuda(), fvs.cuda(), mks.cuda()); torch.cuda.synchronize(); print('racecheck forward ok', float(out.abs().mean()))
" > $O/${TAG}_sanitizer_racecheck.log 2>&1; echo "racecheck rc $?"; tail -4 $O/${TAG}_sanitizer_racecheck.log
ls -la $O | grep ${TAG} | head -40; du -sh $O
