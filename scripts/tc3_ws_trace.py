"""Per-role clock64 timeline of one interior conv_tc3_ws CTA on real-size layers (profiling aid)."""
import ctypes as C, sys, torch
sys.path.insert(0, '.')
from crfp_b200 import _lib as L
from crfp_b200.packing import pack_conv_tc3
h = L.lib()
st = C.c_void_p(torch.cuda.current_stream().cuda_stream)
def run(cin_list, cout, H=360, W=640, act=1, residual=False, rows=8):
    srcs = [torch.randn(1, H, W, c, device='cuda') for c in cin_list]
    w = torch.randn(cout, sum(cin_list), 3, 3, device='cuda') * 0.05
    b = torch.zeros(cout, device='cuda')
    hi, lo, bp, _ = pack_conv_tc3(w, b, cin_list)
    out = torch.empty(1, H, W, cout, device='cuda')
    res = torch.randn(1, H, W, cout, device='cuda')
    d = L.ConvTc3Desc()
    d.n, d.h, d.w, d.nsrc = 1, H, W, len(srcs)
    for i, s in enumerate(srcs):
        d.src[i] = L.TcSrc(ptr=s.data_ptr(), c=s.shape[-1], cstride=s.shape[-1], coffset=0)
    d.cout, d.act = cout, act
    d.weight_hi, d.weight_lo, d.bias = hi.data_ptr(), lo.data_ptr(), bp.data_ptr()
    d.post_scale = 1.0
    d.out_kind, d.ndst = L.TC_OUT_F32, 1
    d.dst[0] = L.TcSrc(ptr=out.data_ptr(), c=cout, cstride=cout, coffset=0)
    if residual:
        d.residual, d.res_cstride, d.res_coffset = res.data_ptr(), cout, 0
    flow = torch.randn(1, H, W, 2, device='cuda')
    if act == 3:
        d.flow, d.head_split, d.head_mag = flow.data_ptr(), 144, 10.0
    flush = torch.empty(64 * 1024 * 1024, device='cuda')
    for _ in range(3):
        L.check(h.crfp_conv3x3_tc3_fwd(C.byref(d), st))
    torch.cuda.synchronize()
    ts = []
    for _ in range(5):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); L.check(h.crfp_conv3x3_tc3_fwd(C.byref(d), st)); e1.record(); torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1) * 1e3)
    tr = torch.zeros(4096, dtype=torch.int64, device='cuda')
    flush.zero_()
    L.check(h.crfp_conv3x3_tc3_trace(C.byref(d), tr.data_ptr(), st)); torch.cuda.synchronize()
    t = tr.cpu()
    t0 = int(t[1])
    print(f"== {cin_list}->{cout} res={residual}: {min(ts):.1f} us/launch (L2 flushed); prologue {t0 - int(t[0])} cycles")
    P, M, E = t[1024:2048].view(-1, 4), t[2048:3072].view(-1, 4), t[3072:4096].view(-1, 4)
    r = lambda x: int(x) - t0 if int(x) else -1
    last = 0
    for i in range(rows + 2):
        print(f"  row {i:2d} | prod: top {r(P[i,0]):6d} waited {r(P[i,1]):6d} done {r(P[i,2]):6d} | mma: top {r(M[i,0]):6d} full {r(M[i,1]):6d} acce {r(M[i,2]):6d} issued {r(M[i,3]):6d} | epi: top {r(E[i,0]):6d} accf {r(E[i,1]):6d} ld {r(E[i,3]):6d} done {r(E[i,2]):6d}")
    nrow = int((E[:60, 2] != 0).sum())
    print(f"  rows {nrow}; last epilogue done at {r(E[nrow-1,2])} cycles -> {r(E[nrow-1,2]) / nrow:.0f} cycles/row")
import os
if os.environ.get('HEADS'): run([32], 216, act=3, rows=12)
else:
    run([32], 32); run([32], 32, residual=True); run([32, 32], 32); run([32], 216, act=0)
