"""Deterministic synthetic rendered pages (SURVEY.md §8 d) — the benchmark/test inputs.

The reference rasterises PDF pages with Poppler (backend/app/pipeline/pdf_extract.py:109-129);
Poppler is not available, so pages are drawn here: white canvas, 1-inch margins, anti-aliased
text lines from a fixed vocabulary, optionally a "photo" block (smooth field + noise) covering
40 % of the text area.  seed = page index; the same seed always yields the same bytes.
"""
from __future__ import annotations

import numpy as np
from PIL import Image, ImageDraw, ImageFont

VOCAB = ("vision compression page image prepare filter deflate base64 kernel batch "
         "memory stream device pixel resize encode document retrieval context answer").split()

PAPER_IN = {"letter": (8.5, 11.0), "a4": (8.27, 11.69), "legal": (8.5, 14.0)}

_font_cache: dict = {}


def _font(px: int):
    f = _font_cache.get(px)
    if f is None:
        f = ImageFont.load_default(size=px)
        _font_cache[px] = f
    return f


def page_size(paper: str = "letter", dpi: int = 200):
    w_in, h_in = PAPER_IN[paper]
    return int(round(w_in * dpi)), int(round(h_in * dpi))


def make_page(seed: int, paper: str = "letter", dpi: int = 200, mode: str = "RGB", photo: bool = False,
              size: tuple[int, int] | None = None) -> Image.Image:
    """One synthetic page as a PIL image (mode 'RGB' or 'L')."""
    rng = np.random.default_rng(seed)
    w, h = size if size is not None else page_size(paper, dpi)
    im = Image.new("RGB", (w, h), (255, 255, 255))
    draw = ImageDraw.Draw(im)
    margin = dpi
    fpx = max(6, int(round(10 * dpi / 72)))
    font = _font(fpx)
    pitch = int(round(1.35 * fpx))
    y = margin
    x0, x1 = margin, w - margin
    while y + pitch < h - margin:
        if rng.random() >= 0.12:
            nwords = int(rng.integers(6, 16))
            words = [VOCAB[int(k)] for k in rng.integers(0, len(VOCAB), nwords)]
            line = " ".join(words)
            while line and draw.textlength(line, font=font) > (x1 - x0):
                line = line.rsplit(" ", 1)[0] if " " in line else line[:-1]
            draw.text((x0, y), line, fill=(0, 0, 0), font=font)
        y += pitch
    if photo:
        tw, th = x1 - x0, h - 2 * margin
        bw, bh = int(tw * 0.8), int(th * 0.5)            # 0.8 * 0.5 = 40 % of the text area
        bx = x0 + int(rng.integers(0, max(1, tw - bw)))
        by = margin + int(rng.integers(0, max(1, th - bh)))
        yy, xx = np.mgrid[0:bh, 0:bw].astype(np.float32)
        field = 128.0 + 90.0 * np.sin(6.0 * xx / bw + seed) * np.cos(4.0 * yy / bh)
        block = field[:, :, None] + np.array([0.0, 15.0, -20.0], np.float32)[None, None, :]
        block = block + rng.normal(0.0, 6.0, size=block.shape).astype(np.float32)
        im.paste(Image.fromarray(np.clip(block, 0, 255).astype(np.uint8), "RGB"), (bx, by))
    if mode == "L":
        im = im.convert("L")
    elif mode != "RGB":
        raise ValueError(mode)
    return im


def make_pages(n: int, paper: str = "letter", dpi: int = 200, mode: str = "RGB", photo_every: int = 0,
               first_seed: int = 0):
    """n pages, seeds first_seed..first_seed+n-1; every `photo_every`-th page is photo-heavy (0 = none)."""
    return [make_page(first_seed + i, paper, dpi, mode, photo=bool(photo_every) and (i % photo_every == photo_every - 1))
            for i in range(n)]


def mixed_page_types():
    """The 48 page types of config C5 (paper x dpi x mode x content), shuffled with seed 0."""
    types = [(p, d, m, c) for p in ("a4", "letter", "legal") for d in (150, 200, 300, 600)
             for m in ("L", "RGB") for c in (False, True)]
    order = np.random.default_rng(0).permutation(len(types))
    return [types[int(i)] for i in order]


# ---------------------------------------------------------------------------------------------- parallel generation
def _gen_one(spec):
    """Worker of PageFactory: one page as raw bytes (size, mode, bytes)."""
    seed, paper, dpi, mode, photo = spec
    im = make_page(seed, paper, dpi, mode, photo)
    return im.size, im.mode, im.tobytes()


class PageFactory:
    """Draws pages on `workers` SPAWNED processes: FreeType rendering holds the GIL (threads do not help), and spawn (not fork)
    keeps it safe to use from a process that already initialised CUDA.  specs = (seed, paper, dpi, mode, photo) tuples."""

    def __init__(self, workers: int):
        import multiprocessing as mp
        from concurrent.futures import ProcessPoolExecutor
        self.ex = ProcessPoolExecutor(max(1, workers), mp_context=mp.get_context("spawn"))

    def arrays(self, specs, out=None):
        """specs -> list of (H, W, 3) / (H, W) uint8 arrays in order; `out` = optional preallocated arrays (e.g. pinned) to fill."""
        res = []
        for i, ((w, h), mode, raw) in enumerate(self.ex.map(_gen_one, specs, chunksize=2)):
            a = np.frombuffer(raw, np.uint8).reshape((h, w, 3) if mode == "RGB" else (h, w))
            if out is not None:
                out[i][...] = a
                a = out[i]
            res.append(a)
        return res

    def images(self, specs):
        return [Image.fromarray(a, "RGB" if a.ndim == 3 else "L") for a in self.arrays(specs)]

    def close(self):
        self.ex.shutdown(wait=True, cancel_futures=True)

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()
