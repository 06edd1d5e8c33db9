"""Stock PyTorch training step on the B200: the oracle port of the reference's PyTorch path (cuDNN convs, ATen
grid_sample, torchvision deform_conv2d — all with their own autograd) + Charbonnier + torch.optim.Adam, TF32 off and
on.  Informational: the "reference GPU training step" our hand-written backward kernels are measured against
(under tests/ because it runs the oracle).  usage: python tests/tools/torch_gpu_train_baseline.py [v7|crop]"""
import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from crfp_b200.synthetic import make_clip, make_state_dict
from test_gpu_zz_training import _oracle_forward_with_grad, _charbonnier

shape = sys.argv[1] if len(sys.argv) > 1 else "v7"
n, t, h, w, fv = {"v7": (1, 7, 64, 112, 128), "crop": (8, 15, 32, 32, 128)}[shape]
params = {k: torch.nn.Parameter(v.cuda()) for k, v in make_state_dict(seed=1).items()}
opt = torch.optim.Adam([{"params": [p for k, p in params.items() if "spynet" not in k], "lr": 2e-4},
                        {"params": [p for k, p in params.items() if "spynet" in k], "lr": 2.5e-5}], betas=(0.9, 0.999), eps=1e-12)
lrs, fvs, mks, _ = make_clip(seed=2, n=n, t=t, h=h, w=w, fv_size=fv)
hr = torch.rand(n, t, 3, 8 * h, 8 * w, generator=torch.Generator().manual_seed(3)).cuda()
lrs, fvs, mks = lrs.cuda(), fvs.cuda(), mks.cuda()


def step():
    loss = _charbonnier(_oracle_forward_with_grad(params, lrs, fvs, mks), hr)
    opt.zero_grad()
    loss.backward()
    opt.step()
    return loss


for tf32 in (False, True):
    torch.backends.cudnn.allow_tf32 = tf32
    torch.backends.cuda.matmul.allow_tf32 = tf32
    for _ in range(3):
        step()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(3):
        step()
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 3
    print(f"stock PyTorch training step on B200 (tf32={tf32}): {ms:.1f} ms/step, {n * t / ms * 1e3:.1f} frames/s, shape {shape}")
