"""Microbenchmark of the HR thin conv (4k -> 4 channels at 1440x2560) through the C ABI (host-bound below ~22 us per call).
usage: [CRFP_THIN_NOTMA=1] python scripts/thin_exp.py"""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from crfp_b200 import ops, _lib as L
from crfp_b200.packing import pack_conv
torch.manual_seed(0)
H, W = 1440, 2560
res = []
for nq, res_on in ((1, False), (1, True), (2, False), (3, False)):
    srcs = [torch.randn(1, H, W, 4, device="cuda") for _ in range(nq)]
    rot = [[torch.randn(1, H, W, 4, device="cuda") for _ in range(nq)] for _ in range(4)]   # rotate inputs: > L2
    wt = torch.randn(4, 4 * nq, 3, 3, device="cuda") * 0.1
    b = torch.randn(4, device="cuda")
    packed = pack_conv(wt, b, [4] * nq)
    r = torch.randn(1, H, W, 4, device="cuda") if res_on else None
    for i in range(3):
        ops.conv3x3_nhwc(rot[i % 4], wt, b, act=L.ACT_LRELU, residual=r, packed=packed)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(40):
        ops.conv3x3_nhwc(rot[i % 4], wt, b, act=L.ACT_LRELU, residual=r, packed=packed)
    e1.record()
    torch.cuda.synchronize()
    res.append(f"nq={nq}{'+res' if res_on else ''}: {e0.elapsed_time(e1) / 40 * 1e3:.1f} us")
print(f"NOTMA={os.environ.get('CRFP_THIN_NOTMA', '0')}  " + "  ".join(res))
