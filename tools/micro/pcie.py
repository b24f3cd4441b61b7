import torch, time
x = torch.empty(157286400, dtype=torch.uint8).pin_memory()
d = torch.empty_like(x, device="cuda")
y = torch.empty(31707136, dtype=torch.uint8).pin_memory()
dy = torch.empty_like(y, device="cuda")
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
for _ in range(2):
    d.copy_(x, non_blocking=True); torch.cuda.synchronize()
t=time.perf_counter()
for _ in range(5):
    d.copy_(x, non_blocking=True)
torch.cuda.synchronize(); dt=(time.perf_counter()-t)/5
print("H2D %.2f ms -> %.1f GB/s" % (dt*1e3, x.numel()/dt/1e9))
t=time.perf_counter()
for _ in range(5):
    y.copy_(dy, non_blocking=True)
torch.cuda.synchronize(); dt=(time.perf_counter()-t)/5
print("D2H %.2f ms -> %.1f GB/s" % (dt*1e3, y.numel()/dt/1e9))
t=time.perf_counter()
for _ in range(5):
    with torch.cuda.stream(s1): d.copy_(x, non_blocking=True)
    with torch.cuda.stream(s2): y.copy_(dy, non_blocking=True)
torch.cuda.synchronize(); dt=(time.perf_counter()-t)/5
print("H2D+D2H concurrent %.2f ms" % (dt*1e3))
