"""Dev tool: when do page runs land during an end-to-end prepare_pages call (pinned arrays in), and how long does turning them into bytes take?"""
import sys, time
sys.path.insert(0, ".")
import numpy as np, torch
import vision_compression_project_b200 as V
from vision_compression_project_b200 import synth, _native as N
pages = [synth.make_page(i, "letter", 200) for i in range(64)]
host = torch.empty((64, 2200, 1700, 3), dtype=torch.uint8, pin_memory=True)
for i, im in enumerate(pages):
    host[i] = torch.from_numpy(np.array(im))
arrs = [host[i].numpy() for i in range(64)]
eng = V.PagePrep(0)
for _ in range(3):
    eng.prepare_pages(arrs)
log = []
orig = N.gather_bytes
def traced(ptr, ranges, threads=4):
    t0 = time.perf_counter(); r = orig(ptr, ranges, threads); log.append((t0, time.perf_counter(), len(ranges), sum(l for _, l in ranges))); return r
N.gather_bytes = traced
for rep in range(3):
    log.clear()
    t = time.perf_counter(); out = eng.prepare_pages(arrs); te = time.perf_counter()
    print(f"step {rep}: {1e3*(te-t):.2f} ms total; gathers (start ms, dur ms, pages, KB):", [(round(1e3*(a-t),2), round(1e3*(b-a),2), n, s//1024) for a, b, n, s in log])
    print("   timing", eng.last_timing)
for rep in range(2):
    t = time.perf_counter(); out = eng.prepare_pages(arrs, want_base64=False); te = time.perf_counter()
    print(f"no-b64 step {rep}: {1e3*(te-t):.2f} ms total", eng.last_timing)
