"""Drop-in page-image preparation: PIL / bytes in -> PNG / base64 bytes out, computed on a B200.

Stands where the reference does, per rasterised page (backend/app/pipeline/pdf_extract.py:129-130, :55;
scripts/extract_pdf_with_gemini.py:151-152, :84; scripts/extract_page_with_gemini.py:119-127):

    page_image.save(page_image_path)                 ->  open(path, "wb").write(prepare_page(page_image).png)
    model.generate_content([prompt, page_image])     ->  generate_content([prompt, {"mime_type": "image/png",
                                                                                    "data": prepared.png}])   # or .b64

Keyword names and meaning follow Pillow (`Image.convert`, `Image.thumbnail` size rule, `Image.resize`,
`Image.save(format="PNG", compress_level=, optimize=)`), so the oracle for any call is literally the Pillow
composition in oracle/pillow_path.py.  All pixel and byte work runs in libvcprep.so (hand-written sm_100a
kernels); Python only plans and moves pointers.  There is no CPU fallback: without the library or a GPU the
call raises.
"""
from __future__ import annotations

import ctypes as C
import math
import os
import threading
from dataclasses import dataclass, field
from typing import Any, List, Optional, Sequence, Tuple

import numpy as np

from . import _native as N

try:  # PIL is only needed when the caller passes PIL images
    from PIL import Image as _PILImage
except Exception:  # pragma: no cover
    _PILImage = None

LANCZOS, BILINEAR, BICUBIC, BOX, HAMMING = 1, 2, 3, 4, 5
_MODE_CH = {"L": 1, "LA": 2, "RGB": 3, "RGBA": 4}
_CH_MODE = {1: "L", 2: "LA", 3: "RGB", 4: "RGBA"}


@dataclass
class PreparedPage:
    png: Optional[bytes]
    b64: Optional[bytes]
    size: Tuple[int, int]
    mode: str
    adler32: int = 0
    n_idat: int = 0
    error: Optional[str] = None
    stats: dict = field(default_factory=dict)

    def inline_data(self, base64_text: bool = False) -> dict:
        """The image part for `model.generate_content([prompt, part])` (pdf_extract.py:55): handing the SDK finished PNG bytes
        instead of the PIL image skips the second full image encode it would do per page.
        base64_text=True gives the REST `inline_data` form (base64 string instead of raw bytes)."""
        if self.png is None:
            raise ValueError(self.error or "page was not prepared")
        if base64_text:
            if self.b64 is None:
                raise ValueError("prepared without base64 (want_base64=False)")
            return {"mime_type": "image/png", "data": self.b64.decode("ascii")}
        return {"mime_type": "image/png", "data": self.png}

    def blob(self) -> Tuple[str, bytes]:
        """(mime type, bytes) — the pair SDK blob constructors take."""
        if self.png is None:
            raise ValueError(self.error or "page was not prepared")
        return "image/png", self.png


def thumbnail_size(src: Tuple[int, int], box: Tuple[int, int]) -> Tuple[int, int]:
    """Aspect-preserving target size of Image.thumbnail (PIL/Image.py:2876-2898): never enlarges."""
    w, h = src
    x, y = box
    if x >= w and y >= h:
        return (w, h)

    def round_aspect(number, key):
        return max(min(math.floor(number), math.ceil(number), key=key), 1)

    aspect = w / h
    if x / y >= aspect:
        x = round_aspect(y * aspect, key=lambda n: abs(aspect - n / y))
    else:
        y = round_aspect(x / aspect, key=lambda n: 0 if n == 0 else abs(aspect - x / n))
    return (x, y)


def parse_pnm(data) -> Tuple[int, int, int, int]:
    """Header of a binary PPM/PGM as `pdftoppm` writes it (P6/P5, maxval 255). Returns (w, h, channels, payload offset)."""
    mv = memoryview(data)
    if len(mv) < 7 or bytes(mv[:2]) not in (b"P6", b"P5"):
        raise ValueError("bytes input must be a binary PPM/PGM (P6/P5) or come with raw_shape=(H, W, C)")
    ch = 3 if bytes(mv[:2]) == b"P6" else 1
    pos, vals = 2, []
    n = len(mv)
    while len(vals) < 3:
        while pos < n and mv[pos] in b" \t\r\n":
            pos += 1
        if pos < n and mv[pos] == 0x23:                      # comment line
            while pos < n and mv[pos] != 0x0A:
                pos += 1
            continue
        start = pos
        while pos < n and mv[pos] not in b" \t\r\n":
            pos += 1
        if start == pos:
            raise ValueError("truncated PNM header")
        vals.append(int(bytes(mv[start:pos])))
    pos += 1                                                 # single whitespace after maxval
    w, h, maxval = vals
    if maxval != 255:
        raise ValueError(f"PNM maxval {maxval} not supported (8-bit pages only)")
    if len(mv) - pos < w * h * ch:
        raise ValueError("truncated PNM payload")
    return w, h, ch, pos


def split_pnm_stream(data) -> list:
    """`pdftoppm` writes the pages of a range back to back on one pipe (pdf2image parses that stream the same way,
    backend/app/pipeline/pdf_extract.py:109-122 via convert_from_path).  Returns one zero-copy memoryview per P6/P5 image."""
    mv = memoryview(data)
    out, pos = [], 0
    while pos < len(mv):
        while pos < len(mv) and mv[pos] in b" \t\r\n":
            pos += 1
        if pos >= len(mv):
            break
        w, h, ch, off = parse_pnm(mv[pos:])
        end = pos + off + w * h * ch
        out.append(mv[pos:end])
        pos = end
    return out


class _Source:
    __slots__ = ("keep", "ptr", "w", "h", "c", "stride", "device", "logical_c", "row_ptrs")

    def __init__(self, keep, ptr, w, h, c, stride, device, logical_c=None):
        self.keep, self.ptr, self.w, self.h, self.c, self.stride, self.device = keep, ptr, w, h, c, stride, device
        self.logical_c = c if logical_c is None else logical_c      # channels of the image's mode (RGB is stored RGBX by Pillow)
        self.row_ptrs = None                                        # address of a table of h row pointers (Pillow's multi-block images)


class _ArrowArray(C.Structure):
    pass


_ArrowArray._fields_ = [("length", C.c_int64), ("null_count", C.c_int64), ("offset", C.c_int64), ("n_buffers", C.c_int64),
                        ("n_children", C.c_int64), ("buffers", C.POINTER(C.c_void_p)),
                        ("children", C.POINTER(C.POINTER(_ArrowArray))), ("dictionary", C.c_void_p),
                        ("release", C.c_void_p), ("private_data", C.c_void_p)]
_capsule_ptr = C.pythonapi.PyCapsule_GetPointer
_capsule_ptr.restype = C.c_void_p
_capsule_ptr.argtypes = [C.py_object, C.c_char_p]


def _pil_zero_copy(image):
    """Pillow's own pixel storage, without the ~20 ms/page `tobytes()` pack: Pillow >= 11.2 exports it through the Arrow
    PyCapsule protocol when the image lives in one block (pages up to 16 MB of storage: letter/A4 at 200 DPI).
    RGB/RGBA are 4 bytes per pixel (RGBX / RGBA), L is 1.  Returns (keepalive, address, storage_channels) or None."""
    if image.mode not in ("RGB", "RGBA", "L") or not hasattr(image, "__arrow_c_array__"):
        return None
    try:
        image.load()
        if image.readonly:                               # wraps foreign memory (frombuffer / mmap): Pillow 12.2's exporter crashes on those
            return None
        schema, array = image.__arrow_c_array__()
    except Exception:                                    # multi-block storage, old Pillow, ...
        return None
    arr = C.cast(_capsule_ptr(array, b"arrow_array"), C.POINTER(_ArrowArray)).contents
    n = image.width * image.height
    if image.mode == "L":
        if arr.n_buffers < 2 or arr.length != n or arr.offset != 0:
            return None
        return (image, schema, array), arr.buffers[1], 1
    if arr.n_children != 1 or arr.length != n or arr.offset != 0:
        return None
    child = arr.children[0].contents
    if child.n_buffers < 2 or child.length != 4 * n or child.offset != 0:
        return None
    return (image, schema, array), child.buffers[1], 4


_capsule_name = C.pythonapi.PyCapsule_GetName
_capsule_name.restype = C.c_char_p
_capsule_name.argtypes = [C.py_object]


def _pil_row_table(image):
    """Images Pillow keeps in several blocks (above 16 MB of storage: 300-DPI pages and up) cannot be exported in one piece, but
    libImaging itself addresses every image through a table of row pointers (`char **image` of ImagingMemoryInstance, Imaging.h).
    Its address is read from the core object's capsule; the struct layout is not assumed but recognised — the (bands, xsize, ysize)
    triple must be found, the three row-table pointers must agree with the mode, pixelsize / linesize must fit, and the first pixel
    must read back as getpixel((0, 0)) — otherwise None (the packed-copy path takes over).
    Returns (keepalive, address of the row table, storage channels)."""
    if image.mode not in ("RGB", "RGBA", "L"):
        return None
    try:
        image.load()
        core = image.im
        cap = core.ptr
        if _capsule_name(cap) != b"Pillow Imaging":
            return None
        base = _capsule_ptr(cap, b"Pillow Imaging")
        w, h = image.size
        bands = {"L": 1, "RGB": 3, "RGBA": 4}[image.mode]
        px = 1 if image.mode == "L" else 4
        ints = (C.c_int32 * 16).from_address(base)
        k = next((i for i in range(1, 12) if ints[i] == w and ints[i + 1] == h and ints[i - 1] == bands), None)
        if k is None:
            return None
        off = ((k + 2) * 4 + 7) & ~7                        # palette, image8, image32, image, block, blocks, then pixelsize, linesize
        ptrs = (C.c_void_p * 6).from_address(base + off)
        table = ptrs[3]
        pixelsize, linesize = (C.c_int32 * 2).from_address(base + off + 48)
        if not table or table != (ptrs[1] if px == 1 else ptrs[2]) or (ptrs[2] if px == 1 else ptrs[1]) or pixelsize != px or linesize < w * px:
            return None
        rows = (C.c_void_p * h).from_address(table)
        if not rows[0] or (h > 1 and rows[1] - rows[0] != linesize):
            return None
        first = bytes((C.c_ubyte * px).from_address(rows[0]))
        want = image.getpixel((0, 0))
        want = bytes([want]) if px == 1 else bytes(want) + (b"" if len(want) == 4 else first[3:4])
        if first != want:
            return None
        return (image, core, cap), table, px
    except Exception:
        return None


def _as_source(image: Any, raw_shape) -> _Source:
    if _PILImage is not None and isinstance(image, _PILImage.Image):
        if image.mode not in _MODE_CH:
            raise ValueError(f"unsupported image mode {image.mode!r} (supported: L, LA, RGB, RGBA)")
        zc = _pil_zero_copy(image)
        if zc is not None:
            keep, addr, sc = zc
            return _Source(keep, addr, image.width, image.height, sc, image.width * sc, False, _MODE_CH[image.mode])
        rt = _pil_row_table(image)
        if rt is not None:
            keep, table, sc = rt
            src = _Source(keep, 0, image.width, image.height, sc, image.width * sc, False, _MODE_CH[image.mode])
            src.row_ptrs = table
            return src
        arr = np.asarray(image)                              # packs Pillow's RGBX storage to interleaved bytes
        if not arr.flags.c_contiguous:
            arr = np.ascontiguousarray(arr)
        c = _MODE_CH[image.mode]
        return _Source(arr, arr.ctypes.data, image.width, image.height, c, image.width * c, False)
    try:
        import torch
        if isinstance(image, torch.Tensor):
            t = image
            if t.dtype != torch.uint8 or t.dim() not in (2, 3):
                raise ValueError("tensor input must be uint8 (H, W) or (H, W, C)")
            h, w = int(t.shape[0]), int(t.shape[1])
            c = 1 if t.dim() == 2 else int(t.shape[2])
            if t.dim() == 3 and (t.stride(2) != 1 or t.stride(1) != c):
                t = t.contiguous()
            if t.dim() == 2 and t.stride(1) != 1:
                t = t.contiguous()
            return _Source(t, t.data_ptr(), w, h, c, int(t.stride(0)), t.is_cuda)
    except ImportError:  # pragma: no cover
        pass
    if isinstance(image, np.ndarray):
        a = image
        if a.dtype != np.uint8 or a.ndim not in (2, 3):
            raise ValueError("array input must be uint8 (H, W) or (H, W, C)")
        c = 1 if a.ndim == 2 else a.shape[2]
        ok = a.strides[-1] == 1 and (a.ndim == 2 or a.strides[1] == c)
        if not ok or a.strides[0] < a.shape[1] * c:
            a = np.ascontiguousarray(a)
        return _Source(a, a.ctypes.data, a.shape[1], a.shape[0], c, a.strides[0], False)
    if isinstance(image, (bytes, bytearray, memoryview)):
        mv = memoryview(image)
        if raw_shape is not None:
            h, w, c = raw_shape
            if len(mv) < h * w * c:
                raise ValueError("raw bytes shorter than raw_shape")
            off = 0
        else:
            w, h, c, off = parse_pnm(mv)
        arr = np.frombuffer(mv, np.uint8, count=w * h * c, offset=off)
        return _Source((image, arr), arr.ctypes.data, w, h, c, w * c, False)
    raise ValueError(f"unsupported input type {type(image).__name__}")


class PagePrep:
    """One device context: a libvcprep handle (stream + arena) and pinned output buffers.  Not shared across
    threads concurrently (calls serialise on the handle); `prepare_page(s)` at module level keep one per thread,
    matching the reference's 5-worker pool (pdf_extract.py:313-333)."""

    def __init__(self, device: int = 0):
        import torch
        self._torch = torch
        self.lib = N.load()
        self.device = device
        hp = C.c_void_p()
        N.check(self.lib.vcp_init(device, C.byref(hp)))
        self.handle = hp
        self._out_png = None
        self._out_b64 = None
        self.max_batch_bytes = 1 << 30
        # host threads that fill the returned bytes objects; share the box's cores between the ranks torchrun started
        try:
            ncpu = len(os.sched_getaffinity(0))
        except AttributeError:  # pragma: no cover
            ncpu = os.cpu_count() or 4
        self.copy_threads = max(2, min(8, ncpu // max(1, int(os.environ.get("LOCAL_WORLD_SIZE", "1")))))
        self.launches_total = 0

    def close(self):
        if getattr(self, "handle", None):
            self.lib.vcp_destroy(self.handle)
            self.handle = None

    def __del__(self):  # pragma: no cover
        try:
            self.close()
        except Exception:
            pass

    # ------------------------------------------------------------------ low level
    def stats(self) -> dict:
        s = N.Stats()
        N.check(self.lib.vcp_get_stats(self.handle, C.byref(s)))
        return s.as_dict()

    def output_bound(self, descs, n, opts) -> Tuple[int, int]:
        p, b = C.c_uint64(), C.c_uint64()
        N.check(self.lib.vcp_output_bound(descs, n, C.byref(opts), C.byref(p), C.byref(b)))
        return p.value, b.value

    def run(self, descs, n, opts, out_png_ptr, png_cap, out_b64_ptr, b64_cap):
        """vcp_prepare_batch; returns the ctypes result array."""
        res = (N.PageResult * n)()
        N.check(self.lib.vcp_prepare_batch(self.handle, descs, n, C.byref(opts), out_png_ptr, png_cap,
                                           out_b64_ptr, b64_cap, res))
        self.launches_total += self.stats()["kernel_launches"]
        return res

    def _pinned(self, which: str, nbytes: int):
        cur = getattr(self, which)
        if cur is None or cur.numel() < nbytes:
            cur = self._torch.empty(int(nbytes * 1.25) + 4096, dtype=self._torch.uint8, pin_memory=True)
            setattr(self, which, cur)
        return cur

    # ------------------------------------------------------------------ decode
    def decode_pages(self, pngs: Sequence[bytes], to_device: bool = False) -> list:
        """PNG bytes -> pixels on the GPU (inflate + un-filter): the reverse of prepare_pages, for re-reading a PNG cache or
        checking output at speed.  Returns one (H, W, C) uint8 array per PNG (numpy, or CUDA tensors with to_device=True);
        a PNG that cannot be decoded gives a ValueError instance in its slot."""
        n = len(pngs)
        if n == 0:
            return []
        bufs = [np.frombuffer(p, np.uint8) for p in pngs]
        ptrs = (C.c_void_p * n)(*[b.ctypes.data for b in bufs])
        lens = (C.c_uint64 * n)(*[b.size for b in bufs])
        total = 0
        for b in bufs:                                       # IHDR sits at a fixed place: size the output from it
            if b.size >= 33:
                w = int.from_bytes(b[16:20].tobytes(), "big"); h = int.from_bytes(b[20:24].tobytes(), "big")
                ch = {0: 1, 4: 2, 2: 3, 6: 4}.get(int(b[25]), 4)
                total += ((w * h * ch + 255) // 256) * 256 if 0 < w < (1 << 24) and 0 < h < (1 << 24) else 0
        total += 4096
        if to_device:
            out = self._torch.empty(total, dtype=self._torch.uint8, device=f"cuda:{self.device}")
        else:
            out = self._pinned("_out_png", total)
        res = (N.DecodeResult * n)()
        N.check(self.lib.vcp_png_decode_batch(self.handle, ptrs, lens, n, out.data_ptr(), out.numel(), int(to_device), res))
        result = []
        good = []
        for r in res:
            if r.status != 0:
                result.append(ValueError("PNG rejected by libvcprep (unsupported or corrupt)"))
                continue
            shape = (r.height, r.width, r.channels)
            if to_device:
                result.append(out[r.pix_off:r.pix_off + r.pix_len].view(shape))
            else:
                a = np.empty(shape, np.uint8)
                result.append(a); good.append((r, a))
        if good:                                             # pinned staging -> the caller's arrays, on several host threads (GIL released)
            m = len(good)
            offs = (C.c_uint64 * m)(*[r.pix_off for r, _ in good])
            lns = (C.c_uint64 * m)(*[r.pix_len for r, _ in good])
            dsts = (C.c_void_p * m)(*[a.ctypes.data for _, a in good])
            N.check(self.lib.vcp_host_scatter(out.data_ptr(), offs, lns, dsts, m, self.copy_threads))
        return result

    # ------------------------------------------------------------------ planning
    @staticmethod
    def _plan(src: _Source, size, max_side, mode, resample, reducing_gap) -> N.PageDesc:
        d = N.PageDesc()
        d.src, d.width, d.height, d.channels, d.row_stride = src.ptr, src.w, src.h, src.c, src.stride
        d.row_ptrs = src.row_ptrs
        target = None
        if size is not None:
            target = (int(size[0]), int(size[1]))
            if target[0] <= 0 or target[1] <= 0:
                raise ValueError("height and width must be > 0")
        elif max_side is not None:
            target = thumbnail_size((src.w, src.h), (int(max_side), int(max_side)))
        if target is not None and target != (src.w, src.h):
            d.dst_width, d.dst_height = target
            if reducing_gap is not None:
                if reducing_gap < 1.0:
                    raise ValueError("reducing_gap must be 1.0 or greater")
                d.reduce_x = int(src.w / target[0] / reducing_gap) or 1     # PIL/Image.py:2413-2416
                d.reduce_y = int(src.h / target[1] / reducing_gap) or 1
        return d

    def prepare_pages(self, images: Sequence[Any], *, size=None, max_side=None, mode: Optional[str] = "RGB",
                      resample: int = LANCZOS, reducing_gap: Optional[float] = None, compress_level: int = 6,
                      optimize: bool = False, filter_mode: str = "pillow", want_base64: bool = True,
                      raw_shape=None) -> List[PreparedPage]:
        if mode not in ("RGB", "L", None):
            raise ValueError(f"unsupported output mode {mode!r} (RGB, L or None = keep)")
        if filter_mode not in ("pillow", "all5"):
            raise ValueError("filter_mode must be 'pillow' (ZipEncode.c rule: Avg only with optimize) or 'all5' (all five candidates)")
        optimize = bool(optimize) or filter_mode == "all5"
        if compress_level == -1:
            compress_level = 6
        if not 0 <= compress_level <= 9:
            raise ValueError("compress_level must be -1 or 0..9")
        n = len(images)
        out: List[Optional[PreparedPage]] = [None] * n
        srcs: List[Optional[_Source]] = [None] * n
        for i, im in enumerate(images):
            try:
                srcs[i] = _as_source(im, raw_shape)
            except (ValueError, TypeError) as e:               # a bad page never fails the batch
                out[i] = PreparedPage(None, None, (0, 0), "", error=f"{type(e).__name__}: {e}")
        # consecutive runs of the same residency (host / device), bounded in bytes
        i = 0
        while i < n:
            if srcs[i] is None:
                i += 1
                continue
            def _key(sr):       # residency + output channels (an RGB image in Pillow's RGBX storage kept "as is" is still 3 channels)
                return (sr.device, {"RGB": 3, "L": 1}.get(mode, 3 if (sr.logical_c == 3 and sr.c == 4) else 0))
            j, nbytes = i, 0
            key = _key(srcs[i])
            dev, out_ch = key
            idx = []
            while j < n and (srcs[j] is None or _key(srcs[j]) == key):
                if srcs[j] is not None:
                    sz = srcs[j].w * srcs[j].h * srcs[j].c
                    if idx and nbytes + sz > self.max_batch_bytes:
                        break
                    idx.append(j)
                    nbytes += sz
                j += 1
            self._run_chunk(idx, srcs, out, dev, size, max_side, out_ch, resample, reducing_gap, compress_level,
                            optimize, want_base64)
            i = j
        return out  # type: ignore[return-value]

    def _run_chunk(self, idx, srcs, out, dev, size, max_side, out_ch, resample, reducing_gap, level, optimize, want_b64):
        import time
        t_0 = time.perf_counter()
        opts = N.Opts()
        opts.out_channels = out_ch
        opts.resample, opts.compress_level, opts.optimize = int(resample), int(level), int(bool(optimize))
        opts.want_b64, opts.src_device, opts.dst_device = int(bool(want_b64)), int(bool(dev)), 0
        good, descs_l = [], []
        for k in idx:
            try:
                d = self._plan(srcs[k], size, max_side, None, resample, reducing_gap)
                N.check(self.lib.vcp_check_page(C.byref(d), C.byref(opts)))      # per-page message, page stays out of the batch
                descs_l.append(d)
                good.append(k)
            except (ValueError, MemoryError, RuntimeError) as e:
                out[k] = PreparedPage(None, None, (0, 0), "", error=f"{type(e).__name__}: {e}")
        m = len(good)
        if m == 0:
            return
        descs = (N.PageDesc * m)(*descs_l)
        bound_png, bound_b64 = self.output_bound(descs, m, opts)
        t_1 = time.perf_counter()
        t_bytes = 0.0
        for attempt in (0, 1):
            cap_png = bound_png if attempt else min(bound_png, max(32 << 20, bound_png // 4))
            cap_b64 = bound_b64 if attempt else min(bound_b64, max(44 << 20, bound_b64 // 4))
            bp = self._pinned("_out_png", cap_png)
            bb = self._pinned("_out_b64", cap_b64) if want_b64 else None
            res = (N.PageResult * m)()
            done = {}
            if m <= 2:
                # a page or two is one launch set: the synchronous call, no worker thread to start and join (single-page latency)
                rc = self.lib.vcp_prepare_batch(self.handle, descs, m, C.byref(opts), bp.data_ptr(), bp.numel(),
                                                bb.data_ptr() if bb is not None else None, bb.numel() if bb is not None else 0, res)
                err = N.error_for(rc, N.last_error()) if rc else None
                if err is None:
                    t_b = time.perf_counter()
                    sel = [i for i in range(m) if res[i].status == 0]
                    pngs = N.gather_bytes(bp.data_ptr(), [(res[i].png_off, res[i].png_len) for i in sel], 1)
                    b64s = (N.gather_bytes(bb.data_ptr(), [(res[i].b64_off, res[i].b64_len) for i in sel], 1)
                            if bb is not None else [None] * len(sel))
                    for i, png, b64 in zip(sel, pngs, b64s):
                        done[i] = (png, b64)
                    t_bytes += time.perf_counter() - t_b
                    break
                if attempt or not (isinstance(err, ValueError) and "too small" in str(err)):
                    raise err
                continue
            # streaming batch: the library's worker thread drives the H2D / kernel / D2H pipeline while this thread turns every
            # finished run of pages into Python bytes (vcp_host_scatter, GIL released) — the copies overlap the GPU work
            N.check(self.lib.vcp_batch_begin(self.handle, descs, m, C.byref(opts), bp.data_ptr(), bp.numel(),
                                             bb.data_ptr() if bb is not None else None, bb.numel() if bb is not None else 0, res))
            err = None
            try:
                first, last = C.c_int(), C.c_int()
                while True:
                    rc = self.lib.vcp_batch_next(self.handle, C.byref(first), C.byref(last))
                    if rc <= 0:
                        if rc < 0:
                            err = N.error_for(rc, N.last_error())
                        break
                    t_b = time.perf_counter()
                    sel = [i for i in range(first.value, last.value + 1) if res[i].status == 0]
                    pngs = N.gather_bytes(bp.data_ptr(), [(res[i].png_off, res[i].png_len) for i in sel], self.copy_threads)
                    b64s = (N.gather_bytes(bb.data_ptr(), [(res[i].b64_off, res[i].b64_len) for i in sel], self.copy_threads)
                            if bb is not None else [None] * len(sel))
                    for i, png, b64 in zip(sel, pngs, b64s):
                        done[i] = (png, b64)
                    t_bytes += time.perf_counter() - t_b
            finally:
                rc_end = self.lib.vcp_batch_end(self.handle)
            if err is None and rc_end < 0:
                err = N.error_for(rc_end, N.last_error())
            if err is None:
                break
            if attempt or not (isinstance(err, ValueError) and "too small" in str(err)):
                raise err
        st = self.stats()
        self.launches_total += st["kernel_launches"]
        t_2 = time.perf_counter()
        for i, k in enumerate(good):
            r = res[i]
            if r.status != 0 or i not in done:
                out[k] = PreparedPage(None, None, (0, 0), "", error=f"page rejected by libvcprep (status {r.status})")
                continue
            png, b64 = done[i]
            out[k] = PreparedPage(png, b64, (r.width, r.height), _CH_MODE[r.channels], r.adler32, r.n_idat, None, st)
        t_2 = t_2 - t_bytes
        self.last_timing = {"plan_ms": 1e3 * (t_1 - t_0), "call_ms": 1e3 * (t_2 - t_1), "bytes_ms": 1e3 * t_bytes}


class _EnginePool:
    """Engines (library handles with their arenas and pinned buffers) are expensive to create (~0.1 s of cudaMalloc /
    cudaMallocHost) and cheap to keep.  Callers borrow one for the duration of a call, so the reference's pattern — a fresh
    5-thread pool per request (pdf_extract.py:313-333) — reuses warm engines instead of initialising one per new thread."""

    def __init__(self):
        self._lock = threading.Lock()
        self._free: dict = {}

    def borrow(self, device: int) -> PagePrep:
        with self._lock:
            lst = self._free.setdefault(device, [])
            if lst:
                return lst.pop()
        return PagePrep(device)

    def give_back(self, eng: PagePrep) -> None:
        with self._lock:
            self._free.setdefault(eng.device, []).append(eng)


_pool = _EnginePool()


def prepare_pages(images: Sequence[Any], *, device: int = 0, **kw) -> List[PreparedPage]:
    """Batched fast path: one launch set for all pages.  A page that cannot be processed gets `.error` set and
    `.png is None`; the others are unaffected (the reference collects failed pages the same way,
    pdf_extract.py:342-350)."""
    eng = _pool.borrow(device)
    try:
        return eng.prepare_pages(images, **kw)
    finally:
        _pool.give_back(eng)


def prepare_stream(batches, *, depth: int = 2, device: int = 0, **kw):
    """A long document as a sequence of page batches: yields `prepare_pages(batch)` for every batch, in order, keeping `depth`
    batches in flight on as many engines (one host thread each, GIL released inside the library).  A single call ends with its
    pipeline draining (the last pages' LZ / Huffman / D2H after the last H2D copy, ~3 ms); with the next batch already copying,
    the PCIe link stays busy.  The reference's analogue is its page thread pool (pdf_extract.py:313-350)."""
    from collections import deque
    from concurrent.futures import ThreadPoolExecutor
    depth = max(1, int(depth))
    with ThreadPoolExecutor(depth) as ex:
        pending = deque()
        for b in batches:
            pending.append(ex.submit(prepare_pages, b, device=device, **kw))
            if len(pending) >= depth:
                yield pending.popleft().result()
        while pending:
            yield pending.popleft().result()


def decode_pages(pngs: Sequence[bytes], *, device: int = 0, to_device: bool = False) -> list:
    """GPU PNG decode of a batch (see PagePrep.decode_pages)."""
    eng = _pool.borrow(device)
    try:
        return eng.decode_pages(pngs, to_device=to_device)
    finally:
        _pool.give_back(eng)


class _Combiner:
    """Micro-batcher for callers that arrive one page at a time (SURVEY.md §8 b: `_process_single_page` runs on 5 worker threads and
    each calls the page body once per page, pdf_extract.py:313-333).
    Measured on B200 (tools/quick_threads.py, letter-200 pages): a one-page launch set is latency-bound (the LZ stage waits for its
    heaviest sub-chunk), so independent launch sets on separate engines overlap almost perfectly — 5 threads reach 3.9 x the rate of
    one — and funnelling them through a queue only lowers the number of sets in flight (1140 against 1600 pages/s).  So the first
    DIRECT concurrent callers each run on their own thread and engine; only beyond that (thread pools much wider than the
    reference's) are pages queued and sent out together, which bounds the engines (arenas, pinned buffers) a process creates."""

    WORKERS = 2          # launch sets the queue keeps in flight beside the direct callers
    DIRECT = 8           # callers served on their own thread before queueing starts

    def __init__(self, device: int):
        import queue
        self.device = device
        self.q: "queue.SimpleQueue" = queue.SimpleQueue()
        self.lock = threading.Lock()
        self.active = 0                                   # callers inside prepare_page on this device
        self.threads = [threading.Thread(target=self._work, name=f"vcprep-combiner-{device}-{k}", daemon=True) for k in range(self.WORKERS)]
        for t in self.threads:
            t.start()

    def _work(self):
        import queue
        while True:
            items = [self.q.get()]
            try:
                while len(items) < 64:
                    items.append(self.q.get_nowait())
            except queue.Empty:
                pass
            groups: dict = {}
            for it in items:
                groups.setdefault(it[1], []).append(it)
            for key, its in groups.items():
                try:
                    res = prepare_pages([it[0] for it in its], device=self.device, **dict(key))
                    for it, r in zip(its, res):
                        it[2].append(r); it[3].set()
                except BaseException as e:  # noqa: BLE001 - handed to the callers
                    for it in its:
                        it[2].append(e); it[3].set()

    def submit(self, image, kw) -> "PreparedPage":
        try:
            key = tuple(sorted(kw.items()))
            hash(key)
        except TypeError:                                 # unhashable keyword value: no batching for this call
            return prepare_pages([image], device=self.device, **kw)[0]
        with self.lock:
            self.active += 1
            direct = self.active <= self.DIRECT
        try:
            if direct:                                    # few callers: each runs its page on its own thread and engine, no hand-off
                return prepare_pages([image], device=self.device, **kw)[0]
            box: list = []
            done = threading.Event()
            self.q.put((image, key, box, done))
            done.wait()
            if isinstance(box[0], BaseException):
                raise box[0]
            return box[0]
        finally:
            with self.lock:
                self.active -= 1


_combiners: dict = {}
_combiners_lock = threading.Lock()


def _combiner(device: int) -> _Combiner:
    c = _combiners.get(device)
    if c is None:
        with _combiners_lock:
            c = _combiners.get(device)
            if c is None:
                c = _combiners[device] = _Combiner(device)
    return c


def prepare_page(image: Any, *, device: int = 0, **kw) -> PreparedPage:
    """Single page; raises (ValueError / MemoryError / RuntimeError) like the Pillow calls it replaces, so the
    reference's per-page try/except (pdf_extract.py:133-136) keeps working.  Concurrent callers (the reference's 5-thread pool)
    are coalesced into shared launch sets, see _Combiner; VCP_NO_COMBINE=1 gives every call its own launch set."""
    if os.environ.get("VCP_NO_COMBINE"):
        r = prepare_pages([image], device=device, **kw)[0]
    else:
        r = _combiner(device).submit(image, kw)
    if r.error is not None:
        kind, _, msg = r.error.partition(": ")
        raise {"ValueError": ValueError, "TypeError": TypeError, "MemoryError": MemoryError}.get(kind, RuntimeError)(msg or r.error)
    return r
