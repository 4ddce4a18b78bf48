"""Dev tool: end-to-end pages/s with PIL images in (pageable RGBX storage), for a few host copy-thread counts (VCP_COPY_THREADS)."""
import os, sys, time
sys.path.insert(0, ".")
import vision_compression_project_b200 as V
from vision_compression_project_b200 import synth
pages = [synth.make_page(i, "letter", 200) for i in range(64)]
eng = V.PagePrep(0)
for _ in range(2):
    eng.prepare_pages(pages)
t = time.perf_counter()
for _ in range(5):
    out = eng.prepare_pages(pages)
dt = time.perf_counter() - t
print(f"copy_threads={os.environ.get('VCP_COPY_THREADS', 'default')}: {5 * 64 / dt:.0f} pages/s  last {eng.last_timing} {({k: round(v, 2) for k, v in eng.stats().items() if k.startswith('ms_')})}", flush=True)
