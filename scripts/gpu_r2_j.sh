#!/bin/bash
# round-2 weight-gradient kernels: tests, V7 / crop training bench, kernel table, A/B against round 1's kernels; streaming latency
mkdir -p gpurun_out
TAG=${TAG:-r2j}
(timeout 900 python -m pytest tests/test_gpu_zz_training.py tests/test_gpu_zzz_spynet.py tests/test_gpu_zzz_runtime.py -q --tb=short -x -rxX > gpurun_out/${TAG}_train_tests_gpu.log 2>&1; echo "pytest exit $?" >> gpurun_out/${TAG}_train_tests_gpu.log)
tail -15 gpurun_out/${TAG}_train_tests_gpu.log
timeout 120 python scripts/bench_train.py --shape v7 --steps 5 --warmup 4 --graphs > gpurun_out/${TAG}_train_bench_v7.json 2> gpurun_out/${TAG}_train_err.txt
CRFP_TRAIN_TC=0 timeout 120 python scripts/bench_train.py --shape v7 --steps 5 --warmup 4 --graphs > gpurun_out/${TAG}_train_bench_v7_notc.json 2>> gpurun_out/${TAG}_train_err.txt
timeout 150 python scripts/bench_train.py --shape crop --steps 3 --warmup 4 --graphs > gpurun_out/${TAG}_train_bench_crop.json 2>> gpurun_out/${TAG}_train_err.txt
timeout 120 python scripts/train_kernel_times.py v7 graphs > gpurun_out/${TAG}_train_kernel_times_v7.txt 2>&1
python - << 'PY'
import json, os
t = os.environ.get("TAG", "r2j")
for k in ("v7", "v7_notc", "crop"):
    try:
        d = json.loads(open(f"gpurun_out/{t}_train_bench_{k}.json").read().strip().splitlines()[-1])
        print(k, round(d["value"], 1), "frames/s", round(d["ms_per_step"], 2), "ms/step")
    except Exception as e:
        print(k, "failed", e)
PY
head -24 gpurun_out/${TAG}_train_kernel_times_v7.txt; tail -3 gpurun_out/${TAG}_train_err.txt
timeout 300 python scripts/bench_stream.py > gpurun_out/${TAG}_stream_1080p.json 2>> gpurun_out/${TAG}_train_err.txt; cat gpurun_out/${TAG}_stream_1080p.json
