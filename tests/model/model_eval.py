"""Explore the deflate model's size vs Pillow on the reference's recorded pages (dev tool, needs /root/reference)."""
import ctypes, io, os, sys, time, zlib
import numpy as np
from PIL import Image
HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "..", ".."))
from oracle import restate as R

class Params(ctypes.Structure):
    _fields_ = [(n, ctypes.c_int32) for n in ("bpp", "hash_bits", "ways", "lane_cap", "too_far", "lazy", "cont_min",
                                               "prime_bytes", "capped_wins", "inwin", "cont_maxd", "sub_bytes", "hash2_bytes", "hash2_bits", "noisy_thresh", "noisy_minlen", "noisy_neard", "cost_maxlen", "cost_margin", "cost_warm", "hash2_ways")] + [("block_bytes", ctypes.c_int64)]
class Stats(ctypes.Structure):
    _fields_ = [("tokens", ctypes.c_int64), ("blocks", ctypes.c_int64), ("stored_blocks", ctypes.c_int64)]

def load():
    lib = ctypes.CDLL(os.path.join(HERE, "libdeflate_model.so"))
    lib.dm_deflate_page.restype = ctypes.c_int64
    lib.dm_deflate_page.argtypes = [ctypes.c_void_p, ctypes.c_int64, ctypes.POINTER(Params), ctypes.c_void_p, ctypes.c_void_p, ctypes.POINTER(Stats)]
    return lib

DEFAULT = dict(bpp=3, hash_bits=11, ways=2, lane_cap=64, too_far=32768, lazy=16, cont_min=258, prime_bytes=32768,
               capped_wins=1, inwin=0, cont_maxd=1, noisy_thresh=0, noisy_minlen=6, noisy_neard=0, cost_maxlen=8, cost_margin=0, cost_warm=64, hash2_ways=1, hash2_bytes=6, hash2_bits=11, sub_bytes=32768, block_bytes=512 * 1024)

def deflate(lib, stream: bytes, **kw):
    p = dict(DEFAULT); p.update(kw)
    P = Params(**p)
    src = np.frombuffer(stream, np.uint8)
    pad = np.concatenate([src, np.zeros(64, np.uint8)])
    out = np.zeros(len(stream) + len(stream) // 8 + 4096, np.uint8)
    st = Stats()
    n = lib.dm_deflate_page(pad.ctypes.data, len(stream), ctypes.byref(P), out.ctypes.data, None, ctypes.byref(st))
    return out[:n].tobytes(), st

def filtered_of(path):
    im = Image.open(path); im.load()
    buf = io.BytesIO(); im.save(buf, format="PNG")
    *_, idat, ok = R.png_split(buf.getvalue())
    z = b"".join(idat)
    return zlib.decompress(z), len(z), im

if __name__ == "__main__":
    lib = load()
    pages = sys.argv[1].split(",") if len(sys.argv) > 1 else ["002", "001", "008", "014"]
    variants = {
        "base": {},
    }
    import json
    if len(sys.argv) > 2:
        variants = json.loads(sys.argv[2])
    data = {}
    for pg in pages:
        data[pg] = filtered_of(f"/root/reference/output/pages/page_{pg}.png")
    for name, kw in variants.items():
        row = []
        for pg in pages:
            f, zl, im = data[pg]
            kw2 = dict(kw); kw2.setdefault("bpp", len(im.getbands()))
            t = time.time()
            z, st = deflate(lib, f, **kw2)
            dt = time.time() - t
            assert zlib.decompress(z) == f, (name, pg)
            row.append(f"{pg}:{len(z)}({len(z)/zl:.3f}) tok={st.tokens} st={st.stored_blocks}/{st.blocks}")
        print(f"{name:28s}", " | ".join(row), flush=True)
