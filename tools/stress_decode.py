"""Dev tool: randomized decode stress — many PNGs (Pillow-written at every level / filter mix, this library's at every effort class, odd sizes
and modes, Z_SYNC_FLUSH cuts) in batches, compared with Pillow's decoder."""
import io, sys, zlib
sys.path.insert(0, ".")
import numpy as np
from PIL import Image
import vision_compression_project_b200 as V
from tests import util as U
from tests.test_gpu_decode import _filtered, _png_from_idats
rng = np.random.default_rng(int(sys.argv[1]) if len(sys.argv) > 1 else 0)
ims, pngs = [], []
for t in range(700):
    big = t % 50 == 0
    h = int(rng.integers(1, 1500 if big else 260)); w = int(rng.integers(1, 2000 if big else 330))
    mode = ["L", "RGB", "RGBA", "LA"][int(rng.integers(0, 4))]
    c = {"L": 1, "LA": 2, "RGB": 3, "RGBA": 4}[mode]
    kind = int(rng.integers(0, 5))
    if kind == 0: px = rng.integers(0, 256, (h, w, c), dtype=np.uint8)
    elif kind == 1: px = np.full((h, w, c), int(rng.integers(0, 256)), np.uint8)
    elif kind == 2: px = (np.add.outer(np.arange(h) * 3, np.arange(w) * 5)[:, :, None] + np.arange(c) * 17).astype(np.uint8)
    elif kind == 3: px = (rng.integers(0, 4, (h, w, c)) * 70).astype(np.uint8)
    else:
        px = np.full((h, w, c), 255, np.uint8); n = max(1, h * w // 40)
        px[rng.integers(0, h, n), rng.integers(0, w, n)] = rng.integers(0, 256, (n, c), dtype=np.uint8)
    im = Image.fromarray(px[:, :, 0] if c == 1 else px, mode)
    ims.append(px)
    enc = int(rng.integers(0, 4))
    if enc == 0:
        pngs.append(U.pillow_png(im, **[{}, {"optimize": True}, {"compress_level": 0}, {"compress_level": 1}, {"compress_level": 9}][int(rng.integers(0, 5))]))
    elif enc == 1:
        pngs.append(V.prepare_pages([im], mode=None, want_base64=False, compress_level=int(rng.integers(0, 10)))[0].png)
    elif enc == 2:
        filt = _filtered(px); co = zlib.compressobj(int(rng.integers(1, 10))); piece = int(rng.integers(50, 70000)); idats = []
        for off in range(0, len(filt), piece):
            idats.append(co.compress(filt[off:off + piece]) + co.flush(zlib.Z_SYNC_FLUSH if rng.integers(0, 2) else zlib.Z_FULL_FLUSH))
        idats[-1] += co.flush()
        pngs.append(_png_from_idats(w, h, c, [d for d in idats if d]))
    else:
        z = zlib.compress(_filtered(px), int(rng.integers(0, 10))); step = int(rng.integers(1, 5000))
        pngs.append(_png_from_idats(w, h, c, [z[i:i + step] for i in range(0, len(z), step)]))
bad = 0
for lo in range(0, len(pngs), 175):
    dec = V.decode_pages(pngs[lo:lo + 175])
    for k, d in enumerate(dec):
        if isinstance(d, Exception) or not np.array_equal(d, ims[lo + k]):
            bad += 1; print("MISMATCH", lo + k, ims[lo + k].shape, type(d))
print("stress decode:", len(pngs), "PNGs,", bad, "bad")
