import torch, time
x = torch.empty(718080000, dtype=torch.uint8, pin_memory=True); x.fill_(7)
d = torch.empty_like(x, device="cuda")
for _ in range(3): d.copy_(x, non_blocking=True)
torch.cuda.synchronize()
t = time.perf_counter()
for _ in range(10): d.copy_(x, non_blocking=True)
torch.cuda.synchronize(); dt = (time.perf_counter() - t) / 10
print(f"one 718 MB H2D: {dt*1e3:.2f} ms = {0.71808/dt:.1f} GB/s")
# 64 separate page copies on 4 streams
pages = [x[i*11220000:(i+1)*11220000] for i in range(64)]
ss = [torch.cuda.Stream() for _ in range(4)]
torch.cuda.synchronize(); t = time.perf_counter()
for r in range(5):
    for i, p in enumerate(pages):
        with torch.cuda.stream(ss[i % 4]):
            d[i*11220000:(i+1)*11220000].copy_(p, non_blocking=True)
torch.cuda.synchronize(); dt = (time.perf_counter() - t) / 5
print(f"64 page copies on 4 streams: {dt*1e3:.2f} ms = {0.71808/dt:.1f} GB/s")
