"""Dev tool: 64-page batches from pageable sources (PIL images, plain numpy arrays), with and without repeated-row elision."""
import os, sys, time
sys.path.insert(0, ".")
import numpy as np
from PIL import Image
import vision_compression_project_b200 as V
from vision_compression_project_b200 import synth
if __name__ == "__main__":
    kind = sys.argv[1] if len(sys.argv) > 1 else "text"
    with synth.PageFactory(12) as fac:
        arrs = fac.arrays([(i, "letter", 200, "RGB", kind == "photo" or (kind == "mix" and i % 4 == 3)) for i in range(64)])
    pil = [Image.fromarray(a, "RGB") for a in arrs]
    for name, src in (("PIL", pil), ("pageable numpy", arrs)):
        for _ in range(2): V.prepare_pages(src)
        t = time.perf_counter()
        for _ in range(6): r = V.prepare_pages(src)
        print(f"{kind} pages, {name}, elision {'off' if os.environ.get('VCP_NO_ELIDE') else 'on'}: {6 * 64 / (time.perf_counter() - t):.0f} pages/s", flush=True)
