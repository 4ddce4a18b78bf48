"""Dev tool: size of the deflate model (== the kernels) vs Pillow on a fixed page set, for parameter variants."""
import json, sys
import numpy as np
from PIL import Image
sys.path.insert(0, ".")
from oracle import restate as R
from tests import model_util as M, util as U
from vision_compression_project_b200 import synth
lib = M.load()
pages = {"Lph->RGB": synth.make_page(3, "letter", 200, "L", True).convert("RGB"),
         "RGBph": synth.make_page(3, "letter", 200, "RGB", True),
         "text": synth.make_page(2, "letter", 200, "RGB", False),
         "ref001": Image.open("tests/golden/ref_page_1.png"),
         "ref014": Image.open("/root/reference/output/pages/page_014.png"),
         "ref008": Image.open("/root/reference/output/pages/page_008.png"),
         "Lph": synth.make_page(3, "letter", 200, "L", True),
         "smooth": Image.fromarray(np.random.default_rng(0).integers(0, 256, (3, 4), dtype=np.uint8), "L").convert("RGB").resize((1700, 2200), 1)}
data = {}
for k, im in pages.items():
    ref = U.pillow_png(im)
    data[k] = (U.png_filtered(ref), sum(len(c) for c in R.png_split(ref)[4]), len(im.getbands()), 1 + im.width * len(im.getbands()))
variants = json.loads(sys.argv[1])
print("variant".ljust(24), *[k.ljust(9) for k in data])
for k, kw in variants.items():
    row = []
    for pg, (filt, zr, bpp, rowlen) in data.items():
        kw2 = dict(kw)
        kw2.setdefault("rowlen", rowlen)          # the kernels probe one filtered row up
        z, st = M.deflate(lib, filt, bpp=bpp, **kw2)
        row.append(f"{len(z)/zr:.3f}".ljust(9))
    print(k.ljust(24), *row, flush=True)
