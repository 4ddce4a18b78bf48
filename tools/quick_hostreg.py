"""Dev tool: what does pinning Pillow's own storage cost?  cudaHostRegister / cudaHostUnregister of letter-200 RGBX-sized buffers
(15 MB), one thread and several, against the packed copy the library does today."""
import sys, time, threading, ctypes
import numpy as np, torch
sys.path.insert(0, ".")
rt = torch.cuda.cudart()
torch.cuda.init()
N = 64
bufs = [np.empty(1700 * 2200 * 4, np.uint8) for _ in range(N)]
for b in bufs: b[::4096] = 1                      # touch every page
def reg(b): return rt.cudaHostRegister(b.ctypes.data, b.nbytes, 0)
def unreg(b): return rt.cudaHostUnregister(b.ctypes.data)
for rep in range(2):
    t = time.perf_counter()
    for b in bufs: assert int(reg(b)) == 0
    t1 = time.perf_counter()
    for b in bufs: assert int(unreg(b)) == 0
    t2 = time.perf_counter()
    print(f"1 thread: register {1e3 * (t1 - t) / N:.3f} ms per 15 MB buffer, unregister {1e3 * (t2 - t1) / N:.3f} ms", flush=True)
for T in (4, 8, 16):
    def work(k):
        for b in bufs[k::T]: reg(b)
        for b in bufs[k::T]: unreg(b)
    ths = [threading.Thread(target=work, args=(k,)) for k in range(T)]
    t = time.perf_counter()
    for th in ths: th.start()
    for th in ths: th.join()
    print(f"{T} threads: register + unregister of {N} buffers {1e3 * (time.perf_counter() - t):.1f} ms wall", flush=True)
# DMA from registered pageable memory vs from cudaHostAlloc memory
dev = torch.empty(bufs[0].nbytes, dtype=torch.uint8, device="cuda")
pin = torch.empty(bufs[0].nbytes, dtype=torch.uint8, pin_memory=True)
reg(bufs[0]); src = torch.from_numpy(bufs[0])
for name, s in (("registered", src), ("cudaHostAlloc", pin)):
    for _ in range(3): dev.copy_(s, non_blocking=True)
    torch.cuda.synchronize(); t = time.perf_counter()
    for _ in range(20): dev.copy_(s, non_blocking=True)
    torch.cuda.synchronize(); dt = (time.perf_counter() - t) / 20
    print(f"H2D from {name} memory: {s.numel() / dt / 1e9:.1f} GB/s", flush=True)
