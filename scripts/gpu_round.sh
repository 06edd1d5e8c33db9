#!/bin/bash
# Round-end style run: both bench arms, launch list, ncu captures of the hot kernels.  Everything under gpurun_out/
# (kept < 64 MiB: the big reports are summarised on the box and deleted).
mkdir -p gpurun_out
TAG=${TAG:-final}
if [ -z "$SKIP_BENCH" ]; then
timeout 900 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/${TAG}_bench_reference.log 2>&1
timeout 1200 python bench.py --steps 3 --warmup 3 > gpurun_out/${TAG}_bench.log 2>&1
tail -c 4000 gpurun_out/${TAG}_bench.log | tail -1 | cut -c1-300
fi
# ncu sees the individual launches: graphs off (a graph replay launches exactly these kernels)
export CRFP_NO_GRAPHS=1
BENCH="python bench.py --frames 3 --steps 1 --warmup 1 --no-cpu-baseline --no-e2e"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/${TAG}_launches_R-lit.csv $BENCH > gpurun_out/ncu_launch.log 2>&1
# DRAM traffic of every launch of one steady-state frame (2 metrics, cheap)
timeout 900 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/${TAG}_dram_R-lit.csv $BENCH > gpurun_out/ncu_dram.log 2>&1
cap() {  # kernel regex, skip, count, keep-report
  timeout 900 ncu --set full --clock-control none --import-source on -k regex:$1 -s $2 -c $3 -f -o gpurun_out/${TAG}_prof_$1 $BENCH > gpurun_out/ncu_$1.log 2>&1
  python scripts/ncu_summary.py gpurun_out/${TAG}_prof_$1.ncu-rep > gpurun_out/${TAG}_$1_ncu_full.txt 2>&1
  python scripts/ncu_hot.py gpurun_out/${TAG}_prof_$1.ncu-rep 30 >> gpurun_out/${TAG}_$1_ncu_full.txt 2>&1
  [ -z "$4" ] && rm -f gpurun_out/${TAG}_prof_$1.ncu-rep
  tail -1 gpurun_out/ncu_$1.log
}
cap conv_tc3_ws_kernel 100 8
cap dcn_tc3_ws_kernel 4 2 keep
cap conv_thin4p_kernel 30 6
cap dcn_hr_kernel 2 1
ls -la gpurun_out | grep ${TAG}; du -sh gpurun_out
