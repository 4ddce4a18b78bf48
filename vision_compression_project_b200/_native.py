"""ctypes binding of libvcprep.so — mirrors include/vcprep.h one to one.  No compute happens in Python."""
from __future__ import annotations

import ctypes as C
import os

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("VCP_LIBRARY") or os.path.join(HERE, "libvcprep.so")    # VCP_LIBRARY: dev A/B builds of the same ABI

VCP_OK, VCP_EINVAL, VCP_ECUDA, VCP_ENOMEM, VCP_ESIZE = 0, -1, -2, -3, -4


class PageDesc(C.Structure):
    _fields_ = [("src", C.c_void_p), ("width", C.c_int32), ("height", C.c_int32), ("channels", C.c_int32),
                ("row_stride", C.c_int64), ("dst_width", C.c_int32), ("dst_height", C.c_int32),
                ("reduce_x", C.c_int32), ("reduce_y", C.c_int32), ("row_ptrs", C.c_void_p)]


class Opts(C.Structure):
    _fields_ = [("out_channels", C.c_int32), ("resample", C.c_int32), ("compress_level", C.c_int32),
                ("optimize", C.c_int32), ("want_b64", C.c_int32), ("src_device", C.c_int32),
                ("dst_device", C.c_int32), ("reserved", C.c_int32)]


class PageResult(C.Structure):
    _fields_ = [("status", C.c_int32), ("width", C.c_int32), ("height", C.c_int32), ("channels", C.c_int32),
                ("png_off", C.c_uint64), ("png_len", C.c_uint64), ("b64_off", C.c_uint64), ("b64_len", C.c_uint64),
                ("adler32", C.c_uint32), ("n_idat", C.c_uint32)]


class DecodeResult(C.Structure):
    _fields_ = [("status", C.c_int32), ("width", C.c_int32), ("height", C.c_int32), ("channels", C.c_int32),
                ("pix_off", C.c_uint64), ("pix_len", C.c_uint64)]


class Stats(C.Structure):
    _fields_ = [(n, C.c_float) for n in ("ms_h2d", "ms_convert", "ms_resample", "ms_filter", "ms_lz", "ms_huff",
                                         "ms_assemble", "ms_b64", "ms_d2h", "ms_total")] + \
               [(n, C.c_uint64) for n in ("kernel_launches", "in_bytes", "filtered_bytes", "png_bytes", "b64_bytes",
                                          "arena_bytes")]

    def as_dict(self):
        return {n: getattr(self, n) for n, _ in self._fields_}


# every symbol include/vcprep.h declares: name -> (restype, argtypes)
SYMBOLS = {
    "vcp_version": (C.c_int, []),
    "vcp_last_error": (C.c_char_p, []),
    "vcp_init": (C.c_int, [C.c_int, C.POINTER(C.c_void_p)]),
    "vcp_destroy": (None, [C.c_void_p]),
    "vcp_check_page": (C.c_int, [C.POINTER(PageDesc), C.POINTER(Opts)]),
    "vcp_output_bound": (C.c_int, [C.POINTER(PageDesc), C.c_int, C.POINTER(Opts), C.POINTER(C.c_uint64), C.POINTER(C.c_uint64)]),
    "vcp_prepare_batch": (C.c_int, [C.c_void_p, C.POINTER(PageDesc), C.c_int, C.POINTER(Opts), C.c_void_p, C.c_uint64,
                                    C.c_void_p, C.c_uint64, C.POINTER(PageResult)]),
    "vcp_batch_begin": (C.c_int, [C.c_void_p, C.POINTER(PageDesc), C.c_int, C.POINTER(Opts), C.c_void_p, C.c_uint64,
                                  C.c_void_p, C.c_uint64, C.POINTER(PageResult)]),
    "vcp_batch_next": (C.c_int, [C.c_void_p, C.POINTER(C.c_int), C.POINTER(C.c_int)]),
    "vcp_batch_end": (C.c_int, [C.c_void_p]),
    "vcp_png_decode_batch": (C.c_int, [C.c_void_p, C.POINTER(C.c_void_p), C.POINTER(C.c_uint64), C.c_int, C.c_void_p, C.c_uint64,
                                       C.c_int, C.POINTER(DecodeResult)]),
    "vcp_get_stats": (C.c_int, [C.c_void_p, C.POINTER(Stats)]),
    "vcp_host_scatter": (C.c_int, [C.c_void_p, C.POINTER(C.c_uint64), C.POINTER(C.c_uint64), C.POINTER(C.c_void_p), C.c_int, C.c_int]),
    "vcp_convert": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int64, C.c_void_p, C.c_int]),
    "vcp_resample_coeffs": (C.c_int, [C.c_int, C.c_int, C.c_int, C.c_float, C.c_float, C.c_void_p, C.c_void_p, C.POINTER(C.c_int)]),
    "vcp_resample": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_int, C.c_int, C.c_int]),
    "vcp_reduce": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_int, C.c_int]),
    "vcp_png_filter": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p, C.POINTER(C.c_uint32)]),
    "vcp_deflate": (C.c_int, [C.c_void_p, C.c_void_p, C.c_uint64, C.c_int, C.c_int, C.c_void_p, C.c_uint64, C.POINTER(C.c_uint64)]),
    "vcp_lz_sub_bytes": (C.c_int, []),
    "vcp_lz_tokens": (C.c_int, [C.c_void_p, C.c_void_p, C.c_uint64, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p]),
    "vcp_adler32": (C.c_int, [C.c_void_p, C.c_void_p, C.c_uint64, C.POINTER(C.c_uint32)]),
    "vcp_crc32": (C.c_int, [C.c_void_p, C.c_void_p, C.c_uint64, C.POINTER(C.c_uint32)]),
    "vcp_base64": (C.c_int, [C.c_void_p, C.c_void_p, C.c_uint64, C.c_void_p]),
}

_lib = None

# CPython C API: a bytes object created with PyBytes_FromStringAndSize(NULL, n) may be filled in place before it is shared
_new_bytes = C.pythonapi.PyBytes_FromStringAndSize
_new_bytes.restype = C.py_object
_new_bytes.argtypes = [C.c_void_p, C.c_ssize_t]
_bytes_ptr = C.pythonapi.PyBytes_AsString
_bytes_ptr.restype = C.c_void_p
_bytes_ptr.argtypes = [C.py_object]


def gather_bytes(src_ptr: int, ranges, threads: int = 4):
    """[(off, len), ...] inside the host buffer at src_ptr -> list of fresh bytes objects (filled by vcp_host_scatter)."""
    n = len(ranges)
    objs = [_new_bytes(None, ln) for _, ln in ranges]
    offs = (C.c_uint64 * n)(*[o for o, _ in ranges])
    lens = (C.c_uint64 * n)(*[ln for _, ln in ranges])
    dsts = (C.c_void_p * n)(*[_bytes_ptr(o) for o in objs])
    check(load().vcp_host_scatter(src_ptr, offs, lens, dsts, n, threads))
    return objs


def load() -> C.CDLL:
    """Load libvcprep.so.  Raises (never falls back) when the CUDA library is missing."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise RuntimeError(
                f"{LIB_PATH} is missing: the CUDA library is the only implementation of this path "
                "(no CPU fallback). Build it with `python -m vision_compression_project_b200.build`.")
        lib = C.CDLL(LIB_PATH)
        for name, (res, args) in SYMBOLS.items():
            fn = getattr(lib, name)
            fn.restype = res
            fn.argtypes = args
        _lib = lib
    return _lib


def last_error() -> str:
    return (load().vcp_last_error() or b"").decode("utf-8", "replace")


def error_for(rc: int, msg: str) -> Exception:
    if rc in (VCP_EINVAL, VCP_ESIZE):
        return ValueError(msg)
    if rc == VCP_ENOMEM:
        return MemoryError(msg)
    return RuntimeError(msg)


def check(rc: int) -> None:
    """Map a VCP_E* code to the exception the reference's per-page try/except expects (pdf_extract.py:133-136)."""
    if rc != VCP_OK:
        raise error_for(rc, last_error())
