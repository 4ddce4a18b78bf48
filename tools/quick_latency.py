"""Dev tool: single-page latency (BASELINE config C1: the reference's recorded page_1.png) and the 5-thread service pattern."""
import sys, time, threading
import numpy as np
sys.path.insert(0, ".")
from PIL import Image
import vision_compression_project_b200 as V
from oracle.pillow_path import prepare_page_cpu
im = Image.open("tests/golden/ref_page_1.png"); im.load()
arr = np.array(im)
for name, src in (("PIL", im), ("numpy", arr)):
    for _ in range(3): V.prepare_page(src)
    t = time.perf_counter(); n = 20
    for _ in range(n): r = V.prepare_page(src)
    print(f"C1 single page from {name}: {(time.perf_counter()-t)/n*1e3:.2f} ms/page, png {len(r.png)} B")
t = time.perf_counter()
for _ in range(3): png, b64, _ = prepare_page_cpu(im)
print(f"C1 Pillow CPU path: {(time.perf_counter()-t)/3*1e3:.1f} ms/page, png {len(png)} B")
def work(k, n):
    for _ in range(n): V.prepare_page(im)
for rep in range(3):        # a fresh 5-thread pool per "request", like extract_pdf_to_page_jsons
    ths = [threading.Thread(target=work, args=(k, 20)) for k in range(5)]
    t = time.perf_counter(); [x.start() for x in ths]; [x.join() for x in ths]
    print(f"request {rep}: 5 threads x 20 pages (pdf_extract.py:313 pattern): {100/(time.perf_counter()-t):.0f} pages/s")
