import ctypes as C, sys, torch
sys.path.insert(0, '.')
from crfp_b200 import _lib as L
h = L.lib()
st = C.c_void_p(torch.cuda.current_stream().cuda_stream)
for ctas in (1, 148, 296):
    for N in (32, 64, 128, 256):
        out = torch.zeros(ctas, dtype=torch.int64, device='cuda')
        for reps in (2000,):
            L.check(h.crfp_selftest_umma_rate(N, reps, ctas, out.data_ptr(), st)); torch.cuda.synchronize()
            L.check(h.crfp_selftest_umma_rate(N, reps, ctas, out.data_ptr(), st)); torch.cuda.synchronize()
            c = out.float()
            print(f"ctas={ctas:4d} N={N:3d}: {c.mean().item()/reps:7.1f} cycles/MMA (max {c.max().item()/reps:.1f})  -> {128*N*16*2/ (c.mean().item()/reps):.0f} FLOP/clk/SM")
