"""NumPy restatement of every stage of the page-image path (TEST INFRASTRUCTURE).

Each function restates, in integer NumPy, the arithmetic of the native routine
Pillow runs when the reference calls ``page_image.save(...)``
(backend/app/pipeline/pdf_extract.py:130) or when the north-star stages
(convert / resize / thumbnail / base64) are applied.  The native sources are not
in /root/reference (Pillow is an un-pinned dependency: requirements.txt:2,
backend/requirements.txt:5; Pillow 12.2.0 is what the image ships), so every
function below is pinned *empirically*: tests/test_oracle.py checks it against
Pillow executed in-process, and the filter restatement additionally against the
reference's recorded PNGs (tests/golden/fixtures.json).

Stage -> native routine it restates:
  convert_mode        libImaging/Convert.c   (L/LA/RGBA/RGB -> RGB, RGB -> L)
  resample_coeffs     libImaging/Resample.c  precompute_coeffs + normalize_coeffs_8bpc
  resample            libImaging/Resample.c  ImagingResampleHorizontal/Vertical_8bpc
  reduce_box          libImaging/Reduce.c
  png_filter          libImaging/ZipEncode.c (adaptive filter selection) + Pack.c
  png_wrap            PIL/PngImagePlugin.py:1325-1525 (_save: magic/IHDR/IDAT/IEND)
  adler32 / crc32     zlib adler32.c / crc32.c
  b64encode           CPython Modules/binascii.c b2a_base64
"""
from __future__ import annotations

import math
import struct
from typing import List, Tuple

import numpy as np

# --------------------------------------------------------------------------- convert

BPP = {"L": 1, "LA": 2, "RGB": 3, "RGBA": 4}


def convert_mode(px: np.ndarray, src_mode: str, dst_mode: str) -> np.ndarray:
    """px: (H, W, C) uint8 with C = BPP[src_mode]. Returns (H, W, BPP[dst_mode])."""
    assert px.dtype == np.uint8 and px.ndim == 3 and px.shape[2] == BPP[src_mode]
    if src_mode == dst_mode:
        return px.copy()
    if dst_mode == "RGB":
        if src_mode in ("L", "LA"):
            return np.repeat(px[:, :, :1], 3, axis=2)          # alpha dropped, no blend
        if src_mode == "RGBA":
            return px[:, :, :3].copy()                         # alpha dropped, no blend
    if dst_mode == "L":
        if src_mode in ("RGB", "RGBA"):
            r = px[:, :, 0].astype(np.uint32)
            g = px[:, :, 1].astype(np.uint32)
            b = px[:, :, 2].astype(np.uint32)
            return (((r * 19595 + g * 38470 + b * 7471 + 0x8000) >> 16).astype(np.uint8))[:, :, None]
        if src_mode == "LA":
            return px[:, :, :1].copy()
    raise ValueError(f"convert {src_mode}->{dst_mode} not on the path")


# --------------------------------------------------------------------------- resample

PRECISION_BITS = 32 - 8 - 2      # 22

NEAREST, LANCZOS, BILINEAR, BICUBIC, BOX, HAMMING = 0, 1, 2, 3, 4, 5


def _sinc(x: float) -> float:
    if x == 0.0:
        return 1.0
    x *= math.pi
    return math.sin(x) / x


def _f_box(x: float) -> float:
    return 1.0 if -0.5 < x <= 0.5 else 0.0


def _f_bilinear(x: float) -> float:
    x = abs(x)
    return 1.0 - x if x < 1.0 else 0.0


def _f_hamming(x: float) -> float:
    x = abs(x)
    if x == 0.0:
        return 1.0
    if x >= 1.0:
        return 0.0
    x *= math.pi
    return math.sin(x) / x * (0.54 + 0.46 * math.cos(x))


def _f_bicubic(x: float) -> float:
    a = -0.5
    x = abs(x)
    if x < 1.0:
        return ((a + 2.0) * x - (a + 3.0)) * x * x + 1
    if x < 2.0:
        return (((x - 5) * x + 8) * x - 4) * a
    return 0.0


def _f_lanczos(x: float) -> float:
    if -3.0 <= x < 3.0:
        return _sinc(x) * _sinc(x / 3)
    return 0.0


FILTERS = {
    BOX: (_f_box, 0.5),
    BILINEAR: (_f_bilinear, 1.0),
    HAMMING: (_f_hamming, 1.0),
    BICUBIC: (_f_bicubic, 2.0),
    LANCZOS: (_f_lanczos, 3.0),
}


def resample_coeffs(in_size: int, out_size: int, flt: int, box: Tuple[float, float] | None = None):
    """Returns (ksize, bounds[out,2] int32 (xmin, n), kk[out,ksize] int32 Q22).

    ``box`` = (in0, in1) is carried as float32 like the C side does.
    """
    fn, fsupport = FILTERS[flt]
    in0 = float(np.float32(0.0 if box is None else box[0]))
    in1 = float(np.float32(in_size if box is None else box[1]))
    scale = (in1 - in0) / out_size
    filterscale = scale if scale >= 1.0 else 1.0
    support = fsupport * filterscale
    ksize = int(math.ceil(support)) * 2 + 1
    bounds = np.zeros((out_size, 2), np.int32)
    kk = np.zeros((out_size, ksize), np.int32)
    ss = 1.0 / filterscale
    for xx in range(out_size):
        center = in0 + (xx + 0.5) * scale
        xmin = int(center - support + 0.5)          # C (int) cast: toward zero
        if xmin < 0:
            xmin = 0
        xmax = int(center + support + 0.5)
        if xmax > in_size:
            xmax = in_size
        n = xmax - xmin
        w = [fn((x + xmin - center + 0.5) * ss) for x in range(n)]
        ww = 0.0
        for v in w:
            ww += v
        if ww != 0.0:
            w = [v / ww for v in w]
        for x, v in enumerate(w):
            t = v * (1 << PRECISION_BITS)
            kk[xx, x] = int(t + 0.5) if v >= 0 else int(t - 0.5)   # int() truncates like the C cast
        bounds[xx] = (xmin, n)
    return ksize, bounds, kk


def _pass_axis1(px: np.ndarray, bounds: np.ndarray, kk: np.ndarray) -> np.ndarray:
    """Resample along axis 1 of (H, W, C) uint8."""
    h, _, c = px.shape
    out_w = bounds.shape[0]
    out = np.empty((h, out_w, c), np.uint8)
    src = px.astype(np.int64)
    for xx in range(out_w):
        xmin, n = int(bounds[xx, 0]), int(bounds[xx, 1])
        acc = np.full((h, c), 1 << (PRECISION_BITS - 1), np.int64)
        if n:
            acc = acc + np.tensordot(src[:, xmin:xmin + n, :], kk[xx, :n].astype(np.int64), axes=([1], [0]))
        out[:, xx, :] = np.clip(acc >> PRECISION_BITS, 0, 255).astype(np.uint8)
    return out


def resample(px: np.ndarray, out_wh: Tuple[int, int], flt: int,
             box: Tuple[float, float, float, float] | None = None) -> np.ndarray:
    """Separable 8-bit resampling: horizontal pass into a uint8 temp, then vertical.

    A pass is skipped when that axis is unchanged (and the box is the full extent).
    """
    h, w, _ = px.shape
    ow, oh = out_wh
    bx = (0.0, 0.0, float(w), float(h)) if box is None else box
    need_h = ow != w or bx[0] != 0.0 or bx[2] != float(w)
    need_v = oh != h or bx[1] != 0.0 or bx[3] != float(h)
    cur = px
    if need_h:
        _, b, k = resample_coeffs(w, ow, flt, (bx[0], bx[2]))
        if need_v:
            # the C code only resamples the rows the vertical pass will read
            _, bv, _ = resample_coeffs(h, oh, flt, (bx[1], bx[3]))
            first = int(bv[0, 0])
            last = int(bv[-1, 0] + bv[-1, 1])
            tmp = _pass_axis1(cur[first:last], b, k)
            _, bv2, kv2 = resample_coeffs(h, oh, flt, (bx[1], bx[3]))
            bv2 = bv2.copy()
            bv2[:, 0] -= first
            out = _pass_axis1(tmp.transpose(1, 0, 2), bv2, kv2).transpose(1, 0, 2)
            return np.ascontiguousarray(out)
        cur = _pass_axis1(cur, b, k)
    if need_v:
        _, bv, kv = resample_coeffs(h, oh, flt, (bx[1], bx[3]))
        cur = np.ascontiguousarray(_pass_axis1(cur.transpose(1, 0, 2), bv, kv).transpose(1, 0, 2))
    return cur.copy() if cur is px else cur


# --------------------------------------------------------------------------- reduce

def reduce_box(px: np.ndarray, fx: int, fy: int) -> np.ndarray:
    """Integer box reduce: ((sum + n/2) * floor(2^24 / n)) >> 24, partial edge cells use their own n."""
    h, w, c = px.shape
    ow, oh = -(-w // fx), -(-h // fy)
    out = np.empty((oh, ow, c), np.uint8)
    src = px.astype(np.uint64)
    for oy in range(oh):
        y0, y1 = oy * fy, min((oy + 1) * fy, h)
        rows = src[y0:y1].sum(axis=0)                      # (W, C)
        for ox in range(ow):
            x0, x1 = ox * fx, min((ox + 1) * fx, w)
            n = (y1 - y0) * (x1 - x0)
            s = rows[x0:x1].sum(axis=0)
            mult = (1 << 24) // n
            out[oy, ox] = (((s + n // 2) * mult) >> 24).astype(np.uint8)
    return out


# --------------------------------------------------------------------------- PNG filter

def _paeth(a: np.ndarray, b: np.ndarray, c: np.ndarray) -> np.ndarray:
    a = a.astype(np.int32); b = b.astype(np.int32); c = c.astype(np.int32)
    p = a + b - c
    pa, pb, pc = np.abs(p - a), np.abs(p - b), np.abs(p - c)
    return np.where((pa <= pb) & (pa <= pc), a, np.where(pb <= pc, b, c))


def _score(res: np.ndarray) -> int:
    v = res.astype(np.int32)
    return int(np.where(v < 128, v, 256 - v).sum())


def filter_row(row: np.ndarray, prev: np.ndarray, bpp: int, optimize: bool = False):
    """One scanline: returns (filter_type, residual bytes). Order/tie rule of ZipEncode.c."""
    n = row.shape[0]
    left = np.zeros(n, np.uint8); left[bpp:] = row[:-bpp] if n > bpp else row[:0]
    ul = np.zeros(n, np.uint8); ul[bpp:] = prev[:-bpp] if n > bpp else prev[:0]
    best_t, best = 0, row.copy()
    best_s = _score(best)
    cands: List[Tuple[int, np.ndarray]] = [
        (2, (row.astype(np.int32) - prev).astype(np.uint8)),
        (1, (row.astype(np.int32) - left).astype(np.uint8)),
    ]
    if optimize:
        cands.append((3, (row.astype(np.int32) - ((left.astype(np.int32) + prev) >> 1)).astype(np.uint8)))
    cands.append((4, (row.astype(np.int32) - _paeth(left, prev, ul)).astype(np.uint8)))
    for t, res in cands:
        if best_s > 0:
            s = _score(res)
            if s < best_s:
                best_t, best, best_s = t, res, s
    return best_t, best


def png_filter(px: np.ndarray, optimize: bool = False) -> np.ndarray:
    """(H, W, C) uint8 -> filtered stream, H * (1 + W*C) bytes, Pillow's adaptive rule.

    Vectorised over rows (the previous row is *pixel* data, so rows are independent);
    ``filter_row`` above is the row-at-a-time statement of the same rule.
    """
    h, w, c = px.shape
    stride = w * c
    cur = px.reshape(h, stride).astype(np.int16)
    up = np.zeros_like(cur); up[1:] = cur[:-1]
    left = np.zeros_like(cur); left[:, c:] = cur[:, :-c]
    ul = np.zeros_like(cur); ul[1:, c:] = cur[:-1, :-c]

    def score(res):
        v = res.astype(np.int32)
        return np.where(v < 128, v, 256 - v).sum(axis=1)

    best = cur.astype(np.uint8)
    best_t = np.zeros(h, np.uint8)
    best_s = score(best)
    cands = [(2, (cur - up).astype(np.uint8)), (1, (cur - left).astype(np.uint8))]
    if optimize:
        cands.append((3, (cur - ((left + up) >> 1)).astype(np.uint8)))
    cands.append((4, (cur - _paeth(left, up, ul)).astype(np.uint8)))
    for t, res in cands:
        s = score(res)
        take = (best_s > 0) & (s < best_s)
        best[take] = res[take]
        best_t[take] = t
        best_s = np.where(take, s, best_s)
    out = np.empty((h, stride + 1), np.uint8)
    out[:, 0] = best_t
    out[:, 1:] = best
    return out.reshape(-1)


def png_unfilter(stream: np.ndarray, h: int, w: int, c: int) -> np.ndarray:
    """Inverse of png_filter for any legal filter bytes (slow, small images only)."""
    stride = w * c
    rows = stream.reshape(h, stride + 1)
    out = np.zeros((h, stride), np.uint8)
    prev = np.zeros(stride, np.int32)
    for y in range(h):
        t = int(rows[y, 0]); r = rows[y, 1:].astype(np.int32)
        cur = np.zeros(stride, np.int32)
        for i in range(stride):
            a = cur[i - c] if i >= c else 0
            b = prev[i]
            cc = prev[i - c] if i >= c else 0
            if t == 0: pred = 0
            elif t == 1: pred = a
            elif t == 2: pred = b
            elif t == 3: pred = (a + b) >> 1
            else:
                p = a + b - cc
                pa, pb, pc = abs(p - a), abs(p - b), abs(p - cc)
                pred = a if (pa <= pb and pa <= pc) else (b if pb <= pc else cc)
            cur[i] = (r[i] + pred) & 255
        out[y] = cur
        prev = cur
    return out.reshape(h, w, c)


# --------------------------------------------------------------------------- checksums, container, base64

def adler32(data: bytes | np.ndarray, start: int = 1) -> int:
    d = np.frombuffer(bytes(data), np.uint8) if not isinstance(data, np.ndarray) else data
    a, b = start & 0xFFFF, (start >> 16) & 0xFFFF
    step = 1 << 16
    for off in range(0, d.size, step):
        blk = d[off:off + step].astype(np.uint64)
        n = blk.size
        s = int(blk.sum())
        wsum = int((blk * np.arange(n, 0, -1, dtype=np.uint64)).sum())
        b = (b + n * a + wsum) % 65521
        a = (a + s) % 65521
    return (b << 16) | a


_CRC_TABLE = None


def _crc_table():
    global _CRC_TABLE
    if _CRC_TABLE is None:
        t = np.zeros(256, np.uint32)
        for n in range(256):
            c = n
            for _ in range(8):
                c = (0xEDB88320 ^ (c >> 1)) if c & 1 else (c >> 1)
            t[n] = c
        _CRC_TABLE = t
    return _CRC_TABLE


def crc32(data: bytes, crc: int = 0) -> int:
    t = _crc_table()
    c = crc ^ 0xFFFFFFFF
    for byte in bytes(data):
        c = int(t[(c ^ byte) & 0xFF]) ^ (c >> 8)
    return c ^ 0xFFFFFFFF


COLOR_TYPE = {1: 0, 2: 4, 3: 2, 4: 6}
PNG_SIG = b"\x89PNG\r\n\x1a\n"


def png_chunk(tag: bytes, data: bytes) -> bytes:
    import zlib
    return struct.pack(">I", len(data)) + tag + data + struct.pack(">I", zlib.crc32(tag + data) & 0xFFFFFFFF)


def png_wrap(w: int, h: int, c: int, zstream: bytes, idat_max: int | None = None) -> bytes:
    """Container as PngImagePlugin._save writes it: sig, IHDR, IDAT x k, IEND."""
    if idat_max is None:
        idat_max = max(65536, w * 4)
    out = [PNG_SIG, png_chunk(b"IHDR", struct.pack(">IIBBBBB", w, h, 8, COLOR_TYPE[c], 0, 0, 0))]
    for off in range(0, len(zstream), idat_max):
        out.append(png_chunk(b"IDAT", zstream[off:off + idat_max]))
    out.append(png_chunk(b"IEND", b""))
    return b"".join(out)


def png_split(png: bytes):
    """Parse a PNG: returns (w, h, bit_depth, color_type, [idat payloads], ok_crc)."""
    import zlib
    assert png[:8] == PNG_SIG, "bad signature"
    off, idat, hdr, ok, seen_end = 8, [], None, True, False
    while off < len(png):
        (ln,) = struct.unpack(">I", png[off:off + 4])
        tag = png[off + 4:off + 8]
        data = png[off + 8:off + 8 + ln]
        (crc,) = struct.unpack(">I", png[off + 8 + ln:off + 12 + ln])
        ok &= (zlib.crc32(tag + data) & 0xFFFFFFFF) == crc
        if tag == b"IHDR":
            hdr = struct.unpack(">IIBBBBB", data)
        elif tag == b"IDAT":
            idat.append(data)
        elif tag == b"IEND":
            seen_end = True
        off += 12 + ln
    assert hdr is not None and seen_end and off == len(png)
    return hdr[0], hdr[1], hdr[2], hdr[3], idat, ok


_B64 = np.frombuffer(b"ABCDEFGHIJKLMNOPQRSTUVWXYZabcdefghijklmnopqrstuvwxyz0123456789+/", np.uint8)


def b64encode(data: bytes) -> bytes:
    d = np.frombuffer(data, np.uint8)
    n = d.size
    pad = (-n) % 3
    d = np.concatenate([d, np.zeros(pad, np.uint8)]).reshape(-1, 3).astype(np.uint32)
    v = (d[:, 0] << 16) | (d[:, 1] << 8) | d[:, 2]
    out = np.stack([_B64[(v >> 18) & 63], _B64[(v >> 12) & 63], _B64[(v >> 6) & 63], _B64[v & 63]], axis=1).reshape(-1)
    if pad:
        out[-pad:] = ord("=")
    return out.tobytes()
