import sys, time
sys.path.insert(0, ".")
import numpy as np
import vision_compression_project_b200 as V
from vision_compression_project_b200 import synth
pages = [synth.make_page(i, "letter", 200) for i in range(8)]
ours = [r.png for r in V.prepare_pages(pages, want_base64=False)]
batch = [ours[i % 8] for i in range(64)]
for _ in range(2): V.decode_pages(batch)
t = time.perf_counter()
for _ in range(3): d = V.decode_pages(batch)
dt = (time.perf_counter() - t) / 3
print(f"decode to host arrays, 64 pages: {64/dt:.0f} pages/s ({dt*1e3:.1f} ms) ok={np.array_equal(d[3], np.asarray(pages[3]))}")
