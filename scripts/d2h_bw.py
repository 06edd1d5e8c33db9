import torch, time
x = torch.empty(1<<30, dtype=torch.uint8, device='cuda')
h = torch.empty(1<<30, dtype=torch.uint8).pin_memory()
for _ in range(2): h.copy_(x, non_blocking=True); torch.cuda.synchronize()
t0=time.perf_counter()
for _ in range(5): h.copy_(x, non_blocking=True)
torch.cuda.synchronize(); dt=time.perf_counter()-t0
print("D2H pinned GB/s", 5*(1<<30)/dt/1e9)
t0=time.perf_counter()
for _ in range(5): x.copy_(h, non_blocking=True)
torch.cuda.synchronize(); dt=time.perf_counter()-t0
print("H2D pinned GB/s", 5*(1<<30)/dt/1e9)
# chunked 44 MB copies on a side stream
s=torch.cuda.Stream()
t0=time.perf_counter()
with torch.cuda.stream(s):
    for i in range(100): h[i*10_000_000:(i*10_000_000)+44_236_800//4].copy_(x[:44_236_800//4], non_blocking=True)
torch.cuda.synchronize(); dt=time.perf_counter()-t0
print("100 x 11 MB chunks GB/s", 100*44_236_800/4/dt/1e9)
