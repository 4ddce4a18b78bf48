"""Dev tool: GPU PNG decode throughput (our PNGs and Pillow's) vs Pillow's decoder on the host."""
import io, sys, time
from concurrent.futures import ThreadPoolExecutor
import numpy as np
sys.path.insert(0, ".")
from PIL import Image
import vision_compression_project_b200 as V
from vision_compression_project_b200 import synth
from tests import util as U
n = 64
kind = sys.argv[1] if len(sys.argv) > 1 else "mix"
pages = [synth.make_page(i, "letter", 200, photo=(kind == "photo") or (kind == "mix" and i % 4 == 3)) for i in range(8)]
ours = [r.png for r in V.prepare_pages(pages, want_base64=False)]
pil = [U.pillow_png(p) for p in pages]
for name, src in (("ours", ours), ("pillow", pil)):
    batch = [src[i % 8] for i in range(n)]
    V.decode_pages(batch[:8])
    for to_dev in (True, False):
        t = time.perf_counter(); d = V.decode_pages(batch, to_device=to_dev); dt = time.perf_counter() - t
        ok = np.array_equal(d[3].cpu().numpy() if to_dev else d[3], np.asarray(pages[3]))
        print(f"decode {name} PNGs, {n} pages, to_device={to_dev}: {n/dt:.0f} pages/s ({dt*1e3:.1f} ms) ok={ok}")
def one(b):
    im = Image.open(io.BytesIO(b)); im.load(); return im.size
batch = [pil[i % 8] for i in range(n)]
with ThreadPoolExecutor(16) as ex:
    t = time.perf_counter(); list(ex.map(one, batch)); dt = time.perf_counter() - t
print(f"Pillow decode on 16 host threads: {n/dt:.0f} pages/s")
