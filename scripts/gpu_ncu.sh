#!/bin/bash
# ncu --set full captures of selected kernels inside a short bench run. usage: KERNELS="a b" SKIP=n bash scripts/gpu_ncu.sh
mkdir -p gpurun_out
BENCH="python bench.py --workload ${WL:-R-lit} --frames 3 --steps 1 --warmup 1 --precision ${PREC:-tc} --no-cpu-baseline --no-e2e"
for K in ${KERNELS:-dcn_tc3_kernel conv_tc3_kernel conv_thin_kernel}; do
  timeout 900 ncu --set full --clock-control none --import-source on -k regex:$K -s ${SKIP:-6} -c ${COUNT:-3} -f -o gpurun_out/prof_$K $BENCH > gpurun_out/ncu_$K.log 2>&1
  tail -2 gpurun_out/ncu_$K.log
done
ls -la gpurun_out/*.ncu-rep
