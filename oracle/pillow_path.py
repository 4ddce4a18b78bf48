"""The reference's CPU path, as the Pillow calls it executes (TEST INFRASTRUCTURE).

The reference saves every rasterised page with ``page_image.save(path)``
(backend/app/pipeline/pdf_extract.py:129-130, scripts/extract_pdf_with_gemini.py:151-152,
scripts/extract_page_with_gemini.py:119-123) and hands the PIL image to the SDK,
which re-encodes it (pdf_extract.py:55).  The north-star pins the full path to

    convert('RGB') -> thumbnail size rule -> resize(LANCZOS) -> PNG -> base64

and this module is that composition, nothing else.  With no keyword arguments it
is byte-for-byte what ``page_image.save(...)`` writes.  It is the checker for the
CUDA path and the timed CPU baseline in bench.py; it is never a fallback.
"""
from __future__ import annotations

import base64
import io
import math
from typing import Optional, Tuple

from PIL import Image


def thumbnail_size(src: Tuple[int, int], box: Tuple[int, int]) -> Tuple[int, int]:
    """Aspect-preserving target size, Pillow's rule (PIL/Image.py:2876-2898).

    Re-derived from Pillow by calling it: ``Image.thumbnail`` on an empty image
    of the given size is the definition; this helper only avoids allocating one.
    """
    w, h = src
    x, y = box
    if x >= w and y >= h:
        return (w, h)

    def round_aspect(number: float, key) -> int:
        return max(min(math.floor(number), math.ceil(number), key=key), 1)

    aspect = w / h
    if x / y >= aspect:
        x = round_aspect(y * aspect, key=lambda n: abs(aspect - n / y))
    else:
        y = round_aspect(x / aspect, key=lambda n: 0 if n == 0 else abs(aspect - x / n))
    return (x, y)


def prepare_page_cpu(
    image: Image.Image,
    *,
    size: Optional[Tuple[int, int]] = None,
    max_side: Optional[int] = None,
    mode: str = "RGB",
    resample: int = Image.Resampling.LANCZOS,
    reducing_gap: Optional[float] = None,
    compress_level: int = -1,
    optimize: bool = False,
    want_base64: bool = True,
):
    """Pillow composition of the hot path. Returns (png_bytes, b64_bytes|None, PIL image)."""
    im = image if image.mode == mode else image.convert(mode)
    if size is None and max_side is not None:
        size = thumbnail_size(im.size, (max_side, max_side))
    if size is not None and tuple(size) != im.size:
        im = im.resize(tuple(size), resample, reducing_gap=reducing_gap)
    buf = io.BytesIO()
    kw = {}
    if compress_level != -1:
        kw["compress_level"] = compress_level
    if optimize:
        kw["optimize"] = True
    im.save(buf, format="PNG", **kw)
    png = buf.getvalue()
    b64 = base64.b64encode(png) if want_base64 else None
    return png, b64, im
