#!/bin/bash
# Training-step artefacts of a round in ONE gpurun call (about 2 GPU-minutes without the ncu pass):
#   tests, bench lines (V7 + training crop, graphed; V7 eager), per-kernel device time, stock-PyTorch baseline,
#   and (NCU=1) an ncu launch list of the backward kernels + a full capture of the weight-gradient kernel.
# usage: TAG=v7 [NCU=1] bash scripts/gpu_train_round.sh      (then copy gpurun_out/${TAG}_train_* to profiles/rNN/)
mkdir -p gpurun_out
TAG=${TAG:-final}
(timeout 300 python -m pytest tests/test_gpu_zz_training.py tests/test_gpu_zzz_spynet.py tests/test_gpu_zzz_runtime.py -q --tb=short -s -rxX \
   > gpurun_out/${TAG}_train_tests_gpu.log 2>&1; echo "pytest exit $?" >> gpurun_out/${TAG}_train_tests_gpu.log)
timeout 120 python scripts/bench_train.py --shape v7 --steps 5 --warmup 4 --graphs > gpurun_out/${TAG}_train_bench_v7.json 2> gpurun_out/${TAG}_train_err.txt
timeout 120 python scripts/bench_train.py --shape v7 --steps 5 --warmup 3 > gpurun_out/${TAG}_train_bench_v7_eager.json 2>> gpurun_out/${TAG}_train_err.txt
timeout 150 python scripts/bench_train.py --shape crop --steps 3 --warmup 4 --graphs > gpurun_out/${TAG}_train_bench_crop.json 2>> gpurun_out/${TAG}_train_err.txt
timeout 120 python scripts/train_kernel_times.py v7 graphs > gpurun_out/${TAG}_train_kernel_times_v7.txt 2>&1
CRFP_WGRAD_THIN=2stage timeout 120 python scripts/train_kernel_times.py v7 graphs > gpurun_out/${TAG}_train_kernel_times_v7_2stage.txt 2>&1
CRFP_WGRAD_THIN=2stage CRFP_WGRAD_KX3=1 timeout 120 python scripts/train_kernel_times.py v7 graphs > gpurun_out/${TAG}_train_kernel_times_v7_2stage_kx3.txt 2>&1
timeout 150 python tests/tools/torch_gpu_train_baseline.py v7 > gpurun_out/${TAG}_train_stock_pytorch_v7.txt 2>&1
if [ -n "$NCU" ]; then
  STEPS=1 timeout 600 ncu --metrics gpu__time_duration.sum,launch__grid_size,sm__throughput.avg.pct_of_peak_sustained_elapsed,l1tex__t_sector_hit_rate.pct,sm__warps_active.avg.pct_of_peak_sustained_active \
    --clock-control none -k regex:"bwd_|conv_wide" --csv --log-file gpurun_out/${TAG}_train_bwd_launches_v7s_ncu.csv python scripts/train_one_step.py v7s > gpurun_out/ncu_train.log 2>&1
  STEPS=1 timeout 600 ncu --set full --clock-control none --import-source on -k regex:conv_bwd_weight_v4 -s 40 -c 4 -f -o gpurun_out/${TAG}_prof_wgrad \
    python scripts/train_one_step.py v7s > gpurun_out/ncu_wgrad.log 2>&1
  python scripts/ncu_summary.py gpurun_out/${TAG}_prof_wgrad.ncu-rep > gpurun_out/${TAG}_train_wgrad_ncu_full.txt 2>&1
  python scripts/ncu_hot.py gpurun_out/${TAG}_prof_wgrad.ncu-rep 30 >> gpurun_out/${TAG}_train_wgrad_ncu_full.txt 2>&1
  rm -f gpurun_out/${TAG}_prof_wgrad.ncu-rep
fi
tail -4 gpurun_out/${TAG}_train_tests_gpu.log; cat gpurun_out/${TAG}_train_bench_v7.json; head -12 gpurun_out/${TAG}_train_kernel_times_v7.txt
