#!/bin/bash
# One gpurun call: headline bench + ncu launch list + ncu full captures of the top kernels.
mkdir -p gpurun_out
WL=${WL:-R-lit}
timeout 900 python bench.py --workload $WL --steps 3 --warmup 3 > gpurun_out/bench_${WL}.log 2>&1
echo "bench exit: $?" >> gpurun_out/bench_${WL}.log
BENCH="python bench.py --workload $WL --frames 3 --steps 1 --warmup 1 --no-cpu-baseline --no-e2e"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_${WL}.csv $BENCH > gpurun_out/ncu_launch.log 2>&1
for K in ${KERNELS:-dcn_l1_kernel conv_wide_kernel conv_thin_kernel dcn_hr_kernel}; do
  timeout 900 ncu --set full --clock-control none --import-source on -k regex:$K -s 4 -c 2 -f -o gpurun_out/prof_$K $BENCH > gpurun_out/ncu_$K.log 2>&1
done
tail -2 gpurun_out/bench_${WL}.log; wc -l gpurun_out/launches_${WL}.csv; ls -la gpurun_out
