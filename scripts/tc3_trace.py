"""Per-phase clock64 trace of one conv_tc3 CTA on a real-size layer (profiling aid)."""
import ctypes as C, sys, torch
sys.path.insert(0, '.')
from crfp_b200 import _lib as L
from crfp_b200.packing import pack_conv_tc3
h = L.lib()
st = C.c_void_p(torch.cuda.current_stream().cuda_stream)
def run(cin_list, cout, H=360, W=640, act=1):
    srcs = [torch.randn(1, H, W, c, device='cuda') for c in cin_list]
    w = torch.randn(cout, sum(cin_list), 3, 3, device='cuda') * 0.05
    b = torch.zeros(cout, device='cuda')
    hi, lo, bp, _ = pack_conv_tc3(w, b, cin_list)
    out = torch.empty(1, H, W, cout, device='cuda')
    d = L.ConvTc3Desc()
    d.n, d.h, d.w, d.nsrc = 1, H, W, len(srcs)
    for i, s in enumerate(srcs):
        d.src[i] = L.TcSrc(ptr=s.data_ptr(), c=s.shape[-1], cstride=s.shape[-1], coffset=0)
    d.cout, d.act = cout, act
    d.weight_hi, d.weight_lo, d.bias = hi.data_ptr(), lo.data_ptr(), bp.data_ptr()
    d.post_scale = 1.0
    d.out_kind, d.ndst = L.TC_OUT_F32, 1
    d.dst[0] = L.TcSrc(ptr=out.data_ptr(), c=cout, cstride=cout, coffset=0)
    tr = torch.zeros(64 * 8, dtype=torch.int64, device='cuda')
    for _ in range(3):
        L.check(h.crfp_conv3x3_tc3_fwd(C.byref(d), st))
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10):
        L.check(h.crfp_conv3x3_tc3_fwd(C.byref(d), st))
    e1.record(); torch.cuda.synchronize()
    L.check(h.crfp_conv3x3_tc3_trace(C.byref(d), tr.data_ptr(), st)); torch.cuda.synchronize()
    t = tr.cpu().view(64, 8)
    print(f"== {cin_list}->{cout}: {e0.elapsed_time(e1)/10*1e3:.1f} us/launch; prologue {int(t[0,1]-t[0,0])} cycles")
    names = ['wait staged', 'convert+sync', 'issue next row', 'issue MMAs', 'wait MMAs', 'epilogue']
    for r in range(1, 9):
        if t[r, 0] == 0: break
        d_ = [int(t[r, k + 1] - t[r, k]) for k in range(6)]
        print(f"   row {r-1}: " + ", ".join(f"{n} {v}" for n, v in zip(names, d_)) + f"  | total {int(t[r,6]-t[r,0])}")
run([32], 32); run([32, 32], 32); run([32], 216, act=0)
