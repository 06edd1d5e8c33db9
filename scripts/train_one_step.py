"""One warm-up + one measured eager training step (for `ncu` launch lists). usage: python scripts/train_one_step.py [v7|crop]"""
import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from crfp_b200 import CRFP_DSV
from crfp_b200.synthetic import make_clip, make_state_dict
from crfp_b200.trainer import Trainer
shape = sys.argv[1] if len(sys.argv) > 1 else "v7"
n, t, h, w, fv = {"v7": (1, 7, 64, 112, 128), "crop": (8, 15, 32, 32, 128), "v7s": (1, 3, 64, 112, 128)}[shape]
model = CRFP_DSV("cuda", mid_channels=32)
model.load_state_dict(make_state_dict(seed=1), strict=True)
model.cuda()
tr = Trainer(model, freeze_flow_iters=0)
lrs, fvs, mks, _ = make_clip(seed=2, n=n, t=t, h=h, w=w, fv_size=fv)
hr = torch.rand(n, t, 3, 8 * h, 8 * w, generator=torch.Generator().manual_seed(3))
batch = (lrs.cuda(), fvs.cuda(), mks.cuda(), hr.cuda())
for _ in range(int(os.environ.get("STEPS", "2"))):
    print("loss", tr.step(*batch).item())
torch.cuda.synchronize()
