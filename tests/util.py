"""Shared helpers of the test-suite (checker side: Pillow, zlib, the oracle)."""
import base64
import hashlib
import io
import os
import zlib

import numpy as np
from PIL import Image

from oracle import restate as R


def sha(b: bytes) -> str:
    return hashlib.sha256(b).hexdigest()[:16]


def pillow_png(im: Image.Image, **kw) -> bytes:
    buf = io.BytesIO()
    im.save(buf, format="PNG", **kw)
    return buf.getvalue()


def png_filtered(png: bytes) -> bytes:
    """Concatenate IDATs, check chunk CRCs, inflate (checks Adler-32)."""
    w, h, bd, ct, idat, ok = R.png_split(png)
    assert ok, "chunk CRC mismatch"
    return zlib.decompress(b"".join(idat))


def decode_all(png: bytes):
    """Decode with every independent decoder available; returns list of (name, ndarray)."""
    out = []
    im = Image.open(io.BytesIO(png)); im.load()
    out.append(("pillow", np.asarray(im)))
    try:
        import cv2
        a = cv2.imdecode(np.frombuffer(png, np.uint8), cv2.IMREAD_UNCHANGED)
        if a is not None:
            if a.ndim == 3 and a.shape[2] == 3:
                a = a[:, :, ::-1]
            elif a.ndim == 3 and a.shape[2] == 4:
                a = a[:, :, [2, 1, 0, 3]]
                if im.mode == "LA":                    # OpenCV expands gray+alpha to BGRA
                    a = a[:, :, [0, 3]]
            out.append(("cv2", a))
    except ImportError:
        pass
    return out


def check_png_against(png: bytes, expect_im: Image.Image, pillow_kw=None, size_tol=1.05):
    """The full contract for one page: valid container, pixels == expect_im, filter bytes == Pillow's,
    size <= size_tol x Pillow's default stream. Returns (our_size, pillow_size)."""
    pillow_kw = pillow_kw or {}
    ref_png = pillow_png(expect_im, **pillow_kw)
    w, h, bd, ct, idat, ok = R.png_split(png)
    assert ok and (w, h) == expect_im.size and bd == 8
    assert ct == {"L": 0, "LA": 4, "RGB": 2, "RGBA": 6}[expect_im.mode]
    z = b"".join(idat)
    assert z[:2] == b"\x78\x9c"
    filt = zlib.decompress(z)
    assert filt == png_filtered(ref_png), "filtered stream differs from Pillow's"
    exp = np.asarray(expect_im)
    for name, a in decode_all(png):
        assert a.shape == exp.shape and np.array_equal(a, exp), f"decoder {name}: pixels differ"
    assert len(png) <= size_tol * len(ref_png) + 64, f"PNG {len(png)} B vs Pillow {len(ref_png)} B"
    return len(png), len(ref_png)


def check_b64(png: bytes, b64: bytes):
    assert b64 == base64.b64encode(png)
    assert base64.b64decode(b64) == png
