"""Dev tool: BASELINE config C4 — a 2,000-page synthetic document at 200 DPI through the public API (host arrays in, bytes out),
on one GPU and sharded by page over all visible GPUs from one process."""
import sys, time
import numpy as np, torch
sys.path.insert(0, ".")
import vision_compression_project_b200 as V
from vision_compression_project_b200 import synth
n = int(sys.argv[1]) if len(sys.argv) > 1 else 2000
uniq = [synth.make_page(i, "letter", 200, photo=(i % 4 == 3)) for i in range(16)]
host = torch.empty((16, 2200, 1700, 3), dtype=torch.uint8, pin_memory=True)
for i, im in enumerate(uniq):
    host[i] = torch.from_numpy(np.array(im))
pages = [host[i % 16].numpy() for i in range(n)]
V.prepare_pages(pages[:64])
t = time.perf_counter(); res = V.prepare_pages(pages); dt = time.perf_counter() - t
assert len(res) == n and all(r.error is None for r in res) and res[16].png == res[0].png
print(f"C4 {n} pages, 1 GPU: {n/dt:.0f} pages/s ({dt:.2f} s), mean PNG {sum(len(r.png) for r in res)/n:.0f} B")
g = torch.cuda.device_count()
if g > 1:
    V.prepare_pages_all_gpus(pages[:64 * g])
    t = time.perf_counter(); res2 = V.prepare_pages_all_gpus(pages); dt = time.perf_counter() - t
    assert [r.png for r in res2[:32]] == [r.png for r in res[:32]]
    print(f"C4 {n} pages, {g} GPUs (one process, threads): {n/dt:.0f} pages/s ({dt:.2f} s)")
