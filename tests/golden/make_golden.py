"""Generate the committed golden vectors from the reference's recorded artefacts and Pillow.

Run in the build container (needs /root/reference):  python tests/golden/make_golden.py

Outputs (all small, all committed):
  fixtures.json    digests of the 23 recorded PNGs in /root/reference/output (pixels, filtered
                   stream = decompressed IDAT, Adler-32, filter-type histogram, container layout,
                   local Pillow re-encode size and its base64 length)
  ref_page_1.png   verbatim copy of /root/reference/output/page_1.png — BASELINE.json configs[0]
                   is literally this page; its decoded pixels are the C1 input
  crops.npz        three 96x80 crops of recorded pages (RGB) + their Pillow outputs:
                   filtered stream, optimize=True filtered stream, L conversion, LANCZOS/BICUBIC/
                   BILINEAR/BOX/HAMMING resizes to (61,47) and (131,117), reduce(2), reduce((3,2))
"""
import base64
import hashlib
import io
import json
import os
import shutil
import struct
import sys
import zlib

import numpy as np
from PIL import Image

HERE = os.path.dirname(os.path.abspath(__file__))
REF = "/root/reference/output"
sys.path.insert(0, os.path.join(HERE, "..", ".."))
from oracle import restate as R  # noqa: E402


def sha(b: bytes) -> str:
    return hashlib.sha256(b).hexdigest()[:16]


def chunks(png: bytes):
    off, out = 8, []
    while off < len(png):
        (ln,) = struct.unpack(">I", png[off:off + 4])
        out.append((png[off + 4:off + 8].decode(), ln))
        off += 12 + ln
    return out


def pillow_filtered(im: Image.Image, **kw) -> bytes:
    buf = io.BytesIO()
    im.save(buf, format="PNG", **kw)
    *_, idat, ok = R.png_split(buf.getvalue())
    assert ok
    return zlib.decompress(b"".join(idat)), buf.getvalue()


def main():
    names = ["page_1.png"] + [f"pages/page_{i:03d}.png" for i in range(1, 23)]
    table = {}
    for n in names:
        raw = open(os.path.join(REF, n), "rb").read()
        im = Image.open(io.BytesIO(raw)); im.load()
        px = im.tobytes()
        w, h, bd, ct, idat, ok = R.png_split(raw)
        filtered = zlib.decompress(b"".join(idat))
        stride = w * 3 + 1
        ftypes = np.frombuffer(filtered, np.uint8)[::stride]
        hist = np.bincount(ftypes, minlength=5).tolist()
        local_filtered, local_png = pillow_filtered(im)
        assert local_filtered == filtered, n          # Pillow-here reproduces the recorded filter choices
        ch = chunks(raw)
        table[n] = {
            "bytes": len(raw), "sha_file": sha(raw), "size": [w, h], "mode": im.mode,
            "bit_depth": bd, "color_type": ct, "crc_ok": bool(ok),
            "sha_px": sha(px), "sha_filtered": sha(filtered), "filtered_bytes": len(filtered),
            "adler32": f"{zlib.adler32(filtered) & 0xFFFFFFFF:08x}",
            "filter_hist_NSUAP": hist,
            "zlib_header": b"".join(idat)[:2].hex(),
            "n_idat": sum(1 for t, _ in ch if t == "IDAT"),
            "idat_max": max(l for t, l in ch if t == "IDAT"),
            "chunk_types": sorted({t for t, _ in ch}),
            "pillow_png_bytes": len(local_png),
            "pillow_png_sha": sha(local_png),
            "pillow_b64_bytes": len(base64.b64encode(local_png)),
        }
        print(n, table[n]["bytes"], table[n]["pillow_png_bytes"], hist)
    json.dump({"pillow": Image.__version__ if hasattr(Image, "__version__") else "12.2.0",
               "zlib": zlib.ZLIB_RUNTIME_VERSION, "fixtures": table},
              open(os.path.join(HERE, "fixtures.json"), "w"), indent=1)
    shutil.copyfile(os.path.join(REF, "page_1.png"), os.path.join(HERE, "ref_page_1.png"))

    # small crops with Pillow-computed outputs
    out = {}
    crops = [("pages/page_014.png", (300, 700)), ("page_1.png", (200, 420)), ("pages/page_008.png", (640, 1000))]
    for ci, (n, (x0, y0)) in enumerate(crops):
        im = Image.open(os.path.join(REF, n)).crop((x0, y0, x0 + 96, y0 + 80))
        out[f"c{ci}_px"] = np.asarray(im)
        out[f"c{ci}_filtered"] = np.frombuffer(pillow_filtered(im)[0], np.uint8)
        out[f"c{ci}_filtered_opt"] = np.frombuffer(pillow_filtered(im, optimize=True)[0], np.uint8)
        g = im.convert("L")
        out[f"c{ci}_L"] = np.asarray(g)
        out[f"c{ci}_L_filtered"] = np.frombuffer(pillow_filtered(g)[0], np.uint8)
        for fname, flt in [("lanczos", 1), ("bilinear", 2), ("bicubic", 3), ("box", 4), ("hamming", 5)]:
            for size in [(61, 47), (131, 117)]:
                out[f"c{ci}_{fname}_{size[0]}x{size[1]}"] = np.asarray(im.resize(size, flt))
        out[f"c{ci}_reduce2"] = np.asarray(im.reduce(2))
        out[f"c{ci}_reduce3x2"] = np.asarray(im.reduce((3, 2)))
    np.savez_compressed(os.path.join(HERE, "crops.npz"), **out)
    print("crops.npz", os.path.getsize(os.path.join(HERE, "crops.npz")))


if __name__ == "__main__":
    main()
