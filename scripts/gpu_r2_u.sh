#!/bin/bash
# coalesced 16x256b epilogue of conv_tc3_ws: parity (bit-identical to the one-pixel-per-thread epilogue) + A/B
mkdir -p gpurun_out
TAG=${TAG:-r2u}
python - << 'PY'
import os, subprocess, sys, torch
sys.path.insert(0, os.getcwd())
code = r"""
import torch, sys
from crfp_b200 import CRFP_DSV
from crfp_b200.synthetic import make_clip, make_state_dict
m = CRFP_DSV('cuda', mid_channels=32).eval(); m.load_state_dict(make_state_dict(seed=1), strict=True); m.cuda(); m.use_graphs = False
lrs, fvs, mks, _ = make_clip(seed=2, n=1, t=4, h=45, w=77, fv_size=96)
out = m(lrs.cuda(), fvs.cuda(), mks.cuda()); torch.save(out.cpu(), sys.argv[1])
"""
for tag, env in (("a", {}), ("b", {"CRFP_TC3_NOFAST16": "1"})):
    subprocess.run([sys.executable, "-c", code, f"/tmp/out_{tag}.pt"], env={**os.environ, **env}, check=True)
a, b = torch.load("/tmp/out_a.pt"), torch.load("/tmp/out_b.pt")
print("fast16 epilogue vs one-pixel-per-thread epilogue: bit-identical =", bool(torch.equal(a, b)), "max abs diff", (a - b).abs().max().item())
PY
(timeout 1500 python -m pytest tests -m gpu -x -q --tb=short > gpurun_out/${TAG}_tests.log 2>&1; echo "pytest exit $?" >> gpurun_out/${TAG}_tests.log)
tail -4 gpurun_out/${TAG}_tests.log
B="python bench.py --steps 4 --warmup 3 --no-cpu-baseline --no-extras --no-e2e"
for v in "" "CRFP_TC3_NOFAST16=1"; do
  env $v timeout 600 $B > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err
  python -c "
import json; d=json.loads(open('gpurun_out/${TAG}_bench.json').read().strip().splitlines()[-1]); print('[$v] value', round(d['value'],1), 'conv frac', round(d['roofline']['frac'],3), 'avg launch us', round(d['roofline']['avg_launch_ms']*1e3,1))"
done
