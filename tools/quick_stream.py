"""Dev tool: end-to-end pages/s of K batches — one synchronous prepare_pages call per batch vs prepare_stream (2 or 3 batches in flight)."""
import sys, time
sys.path.insert(0, ".")
import numpy as np, torch
import vision_compression_project_b200 as V
from vision_compression_project_b200 import synth
pages = [synth.make_page(i, "letter", 200) for i in range(64)]
host = torch.empty((64, 2200, 1700, 3), dtype=torch.uint8, pin_memory=True)
for i, im in enumerate(pages):
    host[i] = torch.from_numpy(np.array(im))
arrs = [host[i].numpy() for i in range(64)]
K = 10
for _ in range(3): ref = V.prepare_pages(arrs)
t = time.perf_counter()
for _ in range(K): out = V.prepare_pages(arrs)
dt = time.perf_counter() - t
print(f"sync: {64*K/dt:.0f} pages/s")
for depth in (2, 3):
    list(V.prepare_stream((arrs for _ in range(4)), depth=depth))
    t = time.perf_counter()
    n = 0
    for out in V.prepare_stream((arrs for _ in range(K)), depth=depth):
        n += len(out)
    dt = time.perf_counter() - t
    print(f"stream depth {depth}: {n/dt:.0f} pages/s  same bytes: {all(a.png == b.png and a.b64 == b.b64 for a, b in zip(out, ref))}")
