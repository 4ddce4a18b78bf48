"""Dev tool: decode 8 Pillow-written PNGs (serial parse path) once, for profiling."""
import sys
sys.path.insert(0, ".")
import numpy as np
import vision_compression_project_b200 as V
from vision_compression_project_b200 import synth
from tests import util as U
pages = [synth.make_page(i, "letter", 200) for i in range(8)]
pil = [U.pillow_png(p) for p in pages]
d = V.decode_pages(pil, to_device=True)
d = V.decode_pages(pil, to_device=True)
print(np.array_equal(d[3].cpu().numpy(), np.asarray(pages[3])))
