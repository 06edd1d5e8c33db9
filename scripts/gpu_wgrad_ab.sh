#!/bin/bash
# A/B of the weight-gradient kernel's launch geometry (CTA size cap, thread target) on one training step.
mkdir -p gpurun_out
for cfg in "128 160000" "128 524288" "64 524288"; do
  set -- $cfg
  echo "== CRFP_WGRAD_BD=$1 CRFP_WGRAD_THREADS=$2" >> gpurun_out/train_wgrad_ab.txt
  CRFP_WGRAD_BD=$1 CRFP_WGRAD_THREADS=$2 timeout 60 python scripts/train_kernel_times.py v7 graphs 2>&1 | grep -E "device span|device busy|conv_bwd_weight" | tail -5 >> gpurun_out/train_wgrad_ab.txt
done
cat gpurun_out/train_wgrad_ab.txt
