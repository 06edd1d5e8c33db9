#!/bin/bash
mkdir -p gpurun_out
TAG=${TAG:-r2o}
(timeout 1500 python -m pytest tests -m gpu -x -q --tb=short > gpurun_out/${TAG}_tests.log 2>&1; echo "pytest exit $?" >> gpurun_out/${TAG}_tests.log)
tail -4 gpurun_out/${TAG}_tests.log
B="python bench.py --steps 4 --warmup 3 --no-cpu-baseline --no-extras"
timeout 600 $B > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err
python - << 'PY'
import json, os
t = os.environ.get("TAG", "r2o")
for k in ("bench",):
    try:
        d = json.loads(open(f"gpurun_out/{t}_{k}.json").read().strip().splitlines()[-1])
        print(k, "value", round(d["value"], 1), "e2e", round(d["e2e"]["value"], 1), "wall", round(d["e2e"]["wall_value"], 1))
    except Exception as e:
        print(k, "failed", e)
PY
CRFP_NO_GRAPHS=1 CRFP_PDL=none timeout 600 python scripts/kernel_times.py --frames 20 --steps 4 2>/dev/null | tail -14
tail -3 gpurun_out/${TAG}_bench.err
