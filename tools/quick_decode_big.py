import sys, io, time
sys.path.insert(0, ".")
import numpy as np
import vision_compression_project_b200 as V
from vision_compression_project_b200 import synth
from tests import util as U
pages = [synth.make_page(1, "legal", 600, photo=True), synth.make_page(2, "a4", 600, mode="L"), synth.make_page(3, "letter", 150)]
ours = [r.png for r in V.prepare_pages(pages, mode=None, want_base64=False)]
pil = [U.pillow_png(p) for p in pages]
for name, src in (("ours", ours), ("pillow", pil)):
    t = time.perf_counter(); d = V.decode_pages(src); dt = time.perf_counter() - t
    print(name, [x.shape for x in d], all(np.array_equal(a, np.asarray(p).reshape(a.shape)) for a, p in zip(d, pages)), f"{dt*1e3:.0f} ms")
