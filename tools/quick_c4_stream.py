"""Dev tool: C4 as bench.py runs it on one rank (prepare_stream over batches of 64, results dropped batch by batch), with the
number of host cores the rank may use limited like at N = 4 / 8 (taskset from outside), to see what binds a rank's rate."""
import sys, time, os
import numpy as np, torch
sys.path.insert(0, ".")
import vision_compression_project_b200 as V
from vision_compression_project_b200 import synth
n = int(sys.argv[1]) if len(sys.argv) > 1 else 500
uniq = [synth.make_page(i, "letter", 200, photo=(i % 4 == 3)) for i in range(16)]
host = torch.empty((16, 2200, 1700, 3), dtype=torch.uint8, pin_memory=True)
for i, im in enumerate(uniq): host[i] = torch.from_numpy(np.array(im))
pages = [host[i % 16].numpy() for i in range(n)]
batches = [pages[i:i + 64] for i in range(0, n, 64)]
for _ in V.prepare_stream(iter(batches[:2]), depth=2): pass
best = 0
for rep in range(2):
    t = time.perf_counter(); k = 0; nb = 0
    for out in V.prepare_stream(iter(batches), depth=2):
        k += len(out); nb += sum(len(r.png) + len(r.b64) for r in out)
    dt = time.perf_counter() - t; best = max(best, k / dt)
print(f"cores {len(os.sched_getaffinity(0))} malloc env {os.environ.get('MALLOC_MMAP_THRESHOLD_', '-')}: C4 stream {n} pages: {best:.0f} pages/s, {nb / k / 1e6:.2f} MB of bytes per page")
