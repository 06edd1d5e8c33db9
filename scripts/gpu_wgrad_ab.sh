mkdir -p gpurun_out
for cfg in "128 160000" "128 600000" "256 300000"; do
  set -- $cfg
  echo "== CRFP_WGRAD_BD=$1 CRFP_WGRAD_THREADS=$2" >> gpurun_out/r01_train_wgrad_ab.txt
  CRFP_WGRAD_BD=$1 CRFP_WGRAD_THREADS=$2 timeout 40 python scripts/train_kernel_times.py v7 graphs 2>&1 | grep -E "device span|device busy|conv_bwd_weight" | tail -4 >> gpurun_out/r01_train_wgrad_ab.txt
done
cat gpurun_out/r01_train_wgrad_ab.txt
