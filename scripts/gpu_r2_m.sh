#!/bin/bash
# K-split tensor-core path for the > 64-channel FNet layers: parity + A/B (clip bench, streaming latency, training step)
mkdir -p gpurun_out
TAG=${TAG:-r2m}
(timeout 1200 python -m pytest tests/test_gpu_model.py tests/test_gpu_long.py tests/test_gpu_half.py tests/test_gpu_tc.py -q --tb=short -x -rxX > gpurun_out/${TAG}_tests.log 2>&1; echo "pytest exit $?" >> gpurun_out/${TAG}_tests.log)
tail -6 gpurun_out/${TAG}_tests.log
B="python bench.py --steps 4 --warmup 3 --no-cpu-baseline --no-extras"
timeout 600 $B > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err
CRFP_NO_KSPLIT=1 timeout 600 $B > gpurun_out/${TAG}_bench_noksplit.json 2>> gpurun_out/${TAG}_bench.err
python - << 'PY'
import json, os
t = os.environ.get("TAG", "r2m")
for k in ("bench", "bench_noksplit"):
    try:
        d = json.loads(open(f"gpurun_out/{t}_{k}.json").read().strip().splitlines()[-1])
        print(k, "value", round(d["value"], 1), "e2e", round(d["e2e"]["value"], 1), "wall", round(d["e2e"]["wall_value"], 1), "conv frac", round(d["roofline"]["frac"], 3))
    except Exception as e:
        print(k, "failed", e)
PY
timeout 300 python scripts/bench_stream.py --modes graph > gpurun_out/${TAG}_stream.json 2>> gpurun_out/${TAG}_bench.err; cat gpurun_out/${TAG}_stream.json
CRFP_NO_KSPLIT=1 timeout 300 python scripts/bench_stream.py --modes graph 2>> gpurun_out/${TAG}_bench.err
timeout 300 python scripts/stream_kernel_times.py 2>/dev/null | head -12
tail -3 gpurun_out/${TAG}_bench.err
