"""Dev tool: a small pass over every kernel for compute-sanitizer (memcheck): odd sizes, all stages, host + device inputs."""
import sys
import numpy as np
import torch
sys.path.insert(0, ".")
import vision_compression_project_b200 as V
from vision_compression_project_b200 import stages as S, synth
rng = np.random.default_rng(0)
ims = [synth.make_page(1, size=(333, 217)), synth.make_page(2, size=(640, 480), photo=True), synth.make_page(3, size=(101, 67), mode="L")]
for im in ims:
    r = V.prepare_page(im, mode=None)
    r = V.prepare_page(im, max_side=64, reducing_gap=2.0)
    r = V.prepare_page(np.array(im), size=(77, 55), resample=V.BICUBIC, optimize=True)
    r = V.prepare_page(torch.from_numpy(np.array(im)).cuda(), mode="L")
res = V.prepare_pages(ims + [np.zeros((3, 5, 4), np.uint8), rng.integers(0, 256, (300, 400, 3), dtype=np.uint8)], mode="RGB")
assert all(x.error is None for x in res)
d = rng.integers(0, 256, 100003, dtype=np.uint8).tobytes()
S.crc32(d); S.adler32(d); S.base64(d); S.deflate(d); S.deflate(bytes(70000)); S.lz_tokens(b"abc" * 20000)
# decode: this library's PNGs (segment chain), Pillow's (serial chain), a multi-IDAT page, odd sizes / modes, a corrupt one
import io
from PIL import Image
big = synth.make_page(4, size=(900, 1300))
pngs = [x.png for x in res] + [V.prepare_page(big).png]
for im in ims + [big]:
    b = io.BytesIO(); im.save(b, format="PNG"); pngs.append(b.getvalue())
bad = bytearray(pngs[-1]); bad[len(bad) // 2:len(bad) // 2 + 32] = bytes(32); pngs.append(bytes(bad))
dec = V.decode_pages(pngs)
assert np.array_equal(dec[len(res)], np.asarray(big))
print("sanitize smoke done")
