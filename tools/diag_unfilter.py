"""Dev tool: where does a decoded page differ from the source?  First mismatches with the row's filter type."""
import sys, zlib, struct
import numpy as np
sys.path.insert(0, ".")
import vision_compression_project_b200 as V
from vision_compression_project_b200 import synth
from PIL import Image
def filt_types(png):
    o = 8; idat = b""; w = h = c = 0
    while o < len(png):
        n, t = struct.unpack(">I4s", png[o:o + 8])
        if t == b"IHDR": w, h, _, ct = struct.unpack(">IIBB", png[o + 8:o + 18]); c = {0: 1, 2: 3, 4: 2, 6: 4}[ct]
        if t == b"IDAT": idat += png[o + 8:o + 8 + n]
        o += 12 + n
    raw = zlib.decompress(idat)
    return [raw[y * (w * c + 1)] for y in range(h)]
rng = np.random.default_rng(0)
cases = {"photo": synth.make_page(1, "letter", 200, photo=True), "L": synth.make_page(2, size=(333, 517), mode="L"),
         "noise": Image.fromarray(rng.integers(0, 256, (300, 400, 3), dtype=np.uint8), "RGB"),
         "text": synth.make_page(3, "letter", 200)}
for name, im in cases.items():
    px = np.asarray(im); px = px[:, :, None] if px.ndim == 2 else px
    png = V.prepare_pages([im], mode=None, want_base64=False)[0].png
    d = V.decode_pages([png])[0]
    d = d[:, :, None] if d.ndim == 2 else d
    bad = np.argwhere(d != px)
    ft = filt_types(png)
    print(name, px.shape, "mismatches", len(bad), "filter histogram", np.bincount(ft, minlength=5).tolist())
    if len(bad):
        rows = np.unique(bad[:, 0])
        print("  bad rows:", rows[:20].tolist(), "... of", len(rows))
        for y, x, ch in bad[:12]:
            print(f"  y={y} (band {y // 32} lane {y % 32} ft {ft[y]}) x={x} ch={ch}: got {d[y, x, ch]} want {px[y, x, ch]}")
        y = bad[0][0]
        xs = bad[bad[:, 0] == y][:, 1]
        print("  first bad row x range", xs.min(), xs.max(), "count", len(xs))
