"""Timing + clock64 timeline of crfp_dcn_align_fused at the R-lit L1 shape (360x640), next to the two-kernel path."""
import ctypes as C, sys, torch
sys.path.insert(0, '.')
from crfp_b200 import _lib as L, ops
from crfp_b200.packing import pack_align_heads, pack_dcn_tc3
h, w = 360, 640
g = torch.Generator().manual_seed(0)
z = torch.randn(1, h, w, 32, generator=g).cuda()
x = torch.randn(1, h, w, 32, generator=g).cuda()
flow = (torch.randn(1, h, w, 2, generator=g) * 2).cuda()
w_off = (torch.randn(144, 32, 3, 3, generator=g) * 0.02).cuda(); b_off = torch.zeros(144).cuda()
w_msk = (torch.randn(72, 32, 3, 3, generator=g) * 0.05).cuda(); b_msk = torch.zeros(72).cuda()
wt = (torch.randn(32, 32, 3, 3, generator=g) * 0.05).cuda(); b = torch.zeros(32).cuda()
wf, bf = pack_align_heads(w_off, b_off, w_msk, b_msk)
hi, lo, bp = pack_dcn_tc3(wt, b, 8)
out = torch.empty(1, h, w, 32, device='cuda')
d = L.AlignFusedDesc(n=1, h=h, w=w, z=z.data_ptr(), z_cstride=32, z_coffset=0, flow=flow.data_ptr(), x=x.data_ptr(), x_cstride=32,
                     x_coffset=0, heads_w=wf.data_ptr(), heads_b=bf.data_ptr(), dcn_w_hi=hi.data_ptr(), dcn_w_lo=lo.data_ptr(),
                     dcn_b=bp.data_ptr(), out=out.data_ptr(), out_cstride=32, out_coffset=0, head_mag=10.0)
st = C.c_void_p(torch.cuda.current_stream().cuda_stream)
flush = torch.empty(64 * 1024 * 1024, device='cuda')
for _ in range(3):
    L.check(L.lib().crfp_dcn_align_fused(C.byref(d), st))
torch.cuda.synchronize()
ts = []
for _ in range(5):
    flush.zero_()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); L.check(L.lib().crfp_dcn_align_fused(C.byref(d), st)); e1.record(); torch.cuda.synchronize()
    ts.append(e0.elapsed_time(e1) * 1e3)
print(f"fused align 360x640: {min(ts):.1f} us/launch (L2 flushed), {sorted(ts)[2]:.1f} median")
tr = torch.zeros(256, dtype=torch.int64, device='cuda')
L.check(L.lib().crfp_dcn_align_fused_trace(C.byref(d), tr.data_ptr(), st)); torch.cuda.synchronize()
t = tr.cpu().view(8, 32)
for i in range(6):
    r = t[i]; t0 = int(r[0])
    q = "  ".join(f"k{k}: raw {int(r[4+3*k])-t0:6d} win {int(r[5+3*k])-t0:6d} done {int(r[6+3*k])-t0:6d}" for k in range(8) if int(r[6+3*k]))
    nxt = int(t[i + 1][0]) - t0
    print(f"round {i}: z {int(r[1])-t0:5d} W0 {int(r[2])-t0:5d} heads {int(r[3])-t0:6d} | {q} | next round {nxt}")
