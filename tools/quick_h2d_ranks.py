"""Dev tool, run under torchrun with N ranks: what does the BOX give when every rank copies at once?
 (i)   copy-only: each rank moves 718 MB (one C2 step) of pinned host memory to its GPU, all ranks together -> aggregate GB/s
 (ii)  the same while each rank's host threads also fill Python bytes objects from a pinned buffer (what vcp_host_scatter does with a
       step's 41 MB of PNG + base64 output), i.e. PCIe traffic plus host memcpy traffic
 (iii) D2H of one step's output (41 MB) alone
Prints one line on rank 0: the e2e ceiling of bench.py at N ranks is 64 pages / (718 MB / per-rank H2D rate)."""
import os, sys, time, threading
import torch, torch.distributed as dist
sys.path.insert(0, ".")
rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(local)
if world > 1: dist.init_process_group("nccl", device_id=torch.device("cuda", local))
def barrier():
    if world > 1: dist.barrier()
    torch.cuda.synchronize()
def allmax(x):
    t = torch.tensor([x], device="cuda", dtype=torch.float64)
    if world > 1: dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())
N = 718080000
x = torch.empty(N, dtype=torch.uint8, pin_memory=True); x.fill_(7)
d = torch.empty_like(x, device="cuda")
out_h = torch.empty(41 << 20, dtype=torch.uint8, pin_memory=True); out_d = torch.empty(41 << 20, dtype=torch.uint8, device="cuda")
def h2d(reps):
    for _ in range(reps): d.copy_(x, non_blocking=True)
    torch.cuda.synchronize()
h2d(2); barrier()
t = time.perf_counter(); h2d(10); dt_copy = allmax(time.perf_counter() - t) / 10
# (ii) with host memcpy threads: bytes objects filled from pinned memory by T threads, as the library's scatter does
from vision_compression_project_b200 import _native as Nn
T = max(2, min(8, len(os.sched_getaffinity(0)) // world))
stop = False
def scatter():
    while not stop:
        Nn.gather_bytes(out_h.data_ptr(), [(i * 650000, 650000) for i in range(64)], T)
th = threading.Thread(target=scatter); th.start()
barrier(); t = time.perf_counter(); h2d(10); dt_busy = allmax(time.perf_counter() - t) / 10
stop = True; th.join()
barrier(); t = time.perf_counter()
for _ in range(20): out_h.copy_(out_d, non_blocking=True)
torch.cuda.synchronize(); dt_d2h = allmax(time.perf_counter() - t) / 20
if rank == 0:
    print(f"ranks {world}: H2D 718 MB/rank all at once: {dt_copy*1e3:.1f} ms -> {N/dt_copy/1e9:.1f} GB/s per rank, {world*N/dt_copy/1e9:.0f} GB/s aggregate "
          f"(ceiling {world*64/dt_copy:.0f} pages/s); with host scatter threads ({T}/rank): {dt_busy*1e3:.1f} ms -> {world*N/dt_busy/1e9:.0f} GB/s aggregate "
          f"(ceiling {world*64/dt_busy:.0f} pages/s); D2H 41 MB: {dt_d2h*1e3:.2f} ms; host cores {len(os.sched_getaffinity(0))}", flush=True)
if world > 1: dist.destroy_process_group()
