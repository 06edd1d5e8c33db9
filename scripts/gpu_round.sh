#!/bin/bash
# Round-end style run: reference arm, headline bench, launch list, ncu full captures.  Everything under gpurun_out/.
mkdir -p gpurun_out
timeout 900 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_reference.log 2>&1
timeout 1200 python bench.py --steps 3 --warmup 3 > gpurun_out/bench_tc.log 2>&1
tail -c 3000 gpurun_out/bench_tc.log | tail -3
BENCH="python bench.py --frames 3 --steps 1 --warmup 1 --no-cpu-baseline --no-e2e"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_tc_R-lit.csv $BENCH > gpurun_out/ncu_launch.log 2>&1
for K in conv_tc3_ws_kernel dcn_tc3_kernel conv_thin4_kernel; do
  timeout 900 ncu --set full --clock-control none --import-source on -k regex:$K -s 6 -c 3 -f -o gpurun_out/prof_$K $BENCH > gpurun_out/ncu_$K.log 2>&1
done
ls -la gpurun_out | head -30
