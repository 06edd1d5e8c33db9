"""Stock PyTorch on the B200: the oracle port of the reference's PyTorch path (cuDNN convs, ATen grid_sample, torchvision
deform_conv2d) moved to CUDA, TF32 off and on.  Informational ("the real reference GPU kernel to beat")."""
import os, sys, time
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from crfp_b200.synthetic import make_clip, make_state_dict
from oracle import crfp_oracle as O
h, w, t = 180, 320, int(os.environ.get("FRAMES", "6"))
sd = {k: v.cuda() for k, v in make_state_dict(seed=1).items()}
lrs, fvs, mks, _ = make_clip(seed=2, n=1, t=t, h=h, w=w, fv_size=96)
lrs, fvs, mks = lrs.cuda(), fvs.cuda(), mks.cuda()
for tf32 in (False, True):
    torch.backends.cudnn.allow_tf32 = tf32
    torch.backends.cuda.matmul.allow_tf32 = tf32
    with torch.no_grad():
        for _ in range(2):
            out = O.crfp_dsv_forward(sd, lrs, fvs, mks)
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for _ in range(3):
            out = O.crfp_dsv_forward(sd, lrs, fvs, mks)
        torch.cuda.synchronize()
        dt = (time.perf_counter() - t0) / 3
    print(f"stock PyTorch on B200 (tf32={tf32}): {t / dt:.2f} frames/s ({dt / t * 1e3:.1f} ms/frame), R-lit, {t}-frame clip")
