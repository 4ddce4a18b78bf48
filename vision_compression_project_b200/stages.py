"""Stage-level entry points of libvcprep (device tensors in / out) — what the parity tests call.

Each function is a thin wrapper over one `vcp_*` stage symbol of include/vcprep.h; torch is used only to own
device memory.  Nothing here computes on the host.
"""
from __future__ import annotations

import ctypes as C

import numpy as np
import torch

from . import _native as N
from .api import PagePrep

_engines: dict = {}


def engine(device: int = 0) -> PagePrep:
    e = _engines.get(device)
    if e is None:
        e = _engines[device] = PagePrep(device)
    return e


def _dev(a, device=0) -> torch.Tensor:
    if isinstance(a, torch.Tensor):
        return a.to(f"cuda:{device}").contiguous()
    return torch.from_numpy(np.ascontiguousarray(a)).to(f"cuda:{device}")


def convert(px, dst_channels: int, device: int = 0) -> torch.Tensor:
    e = engine(device)
    t = _dev(px, device)
    h, w = t.shape[0], t.shape[1]
    sc = 1 if t.dim() == 2 else t.shape[2]
    out = torch.empty((h, w, dst_channels), dtype=torch.uint8, device=t.device)
    N.check(e.lib.vcp_convert(e.handle, t.data_ptr(), w, h, sc, 0, out.data_ptr(), dst_channels))
    return out


def resample(px, out_wh, flt: int, device: int = 0) -> torch.Tensor:
    e = engine(device)
    t = _dev(px, device)
    h, w = t.shape[0], t.shape[1]
    c = 1 if t.dim() == 2 else t.shape[2]
    ow, oh = out_wh
    out = torch.empty((oh, ow, c), dtype=torch.uint8, device=t.device)
    N.check(e.lib.vcp_resample(e.handle, t.data_ptr(), w, h, c, out.data_ptr(), ow, oh, flt))
    return out


def reduce(px, fx: int, fy: int, device: int = 0) -> torch.Tensor:
    e = engine(device)
    t = _dev(px, device)
    h, w = t.shape[0], t.shape[1]
    c = 1 if t.dim() == 2 else t.shape[2]
    out = torch.empty((-(-h // fy), -(-w // fx), c), dtype=torch.uint8, device=t.device)
    N.check(e.lib.vcp_reduce(e.handle, t.data_ptr(), w, h, c, out.data_ptr(), fx, fy))
    return out


def png_filter(px, optimize: bool = False, device: int = 0):
    """Returns (filtered stream uint8 tensor, adler32 of it)."""
    e = engine(device)
    t = _dev(px, device)
    h, w = t.shape[0], t.shape[1]
    c = 1 if t.dim() == 2 else t.shape[2]
    out = torch.empty(h * (1 + w * c), dtype=torch.uint8, device=t.device)
    ad = C.c_uint32()
    N.check(e.lib.vcp_png_filter(e.handle, t.data_ptr(), w, h, c, int(optimize), out.data_ptr(), C.byref(ad)))
    return out, ad.value


def deflate(stream, bpp: int = 3, level: int = 6, device: int = 0) -> bytes:
    e = engine(device)
    t = _dev(np.frombuffer(stream, np.uint8) if isinstance(stream, (bytes, bytearray)) else stream, device).reshape(-1)
    n = t.numel()
    cap = n + n // 1000 * 6 + 4096
    out = torch.empty(cap, dtype=torch.uint8, device=t.device)
    ol = C.c_uint64()
    N.check(e.lib.vcp_deflate(e.handle, t.data_ptr(), n, bpp, level, out.data_ptr(), cap, C.byref(ol)))
    return out[:ol.value].cpu().numpy().tobytes()


def lz_tokens(stream, bpp: int = 3, device: int = 0):
    """Returns (tokens [len] uint32 numpy, sub_ntok [nsub], sub_hist [nsub, 316])."""
    e = engine(device)
    t = _dev(np.frombuffer(stream, np.uint8) if isinstance(stream, (bytes, bytearray)) else stream, device).reshape(-1)
    n = t.numel()
    nblk = -(-n // (512 * 1024))
    nsub = 0
    sub_bytes = e.lib.vcp_lz_sub_bytes()
    for b in range(nblk):
        ln = min(512 * 1024, n - b * 512 * 1024)
        nsub += -(-ln // sub_bytes)
    tok = torch.empty(n, dtype=torch.int32, device=t.device)
    ntok = np.zeros(nsub, np.uint32)
    hist = np.zeros((nsub, 316), np.uint32)
    N.check(e.lib.vcp_lz_tokens(e.handle, t.data_ptr(), n, bpp, tok.data_ptr(), ntok.ctypes.data, hist.ctypes.data))
    return tok.cpu().numpy().view(np.uint32), ntok, hist


def adler32(data, device: int = 0) -> int:
    e = engine(device)
    t = _dev(np.frombuffer(data, np.uint8) if isinstance(data, (bytes, bytearray)) else data, device).reshape(-1)
    o = C.c_uint32()
    N.check(e.lib.vcp_adler32(e.handle, t.data_ptr() if t.numel() else None, t.numel(), C.byref(o)))
    return o.value


def crc32(data, device: int = 0) -> int:
    e = engine(device)
    t = _dev(np.frombuffer(data, np.uint8) if isinstance(data, (bytes, bytearray)) else data, device).reshape(-1)
    o = C.c_uint32()
    N.check(e.lib.vcp_crc32(e.handle, t.data_ptr() if t.numel() else None, t.numel(), C.byref(o)))
    return o.value


def base64(data, device: int = 0) -> bytes:
    e = engine(device)
    t = _dev(np.frombuffer(data, np.uint8) if isinstance(data, (bytes, bytearray)) else data, device).reshape(-1)
    n = t.numel()
    out = torch.empty(4 * ((n + 2) // 3) + 16, dtype=torch.uint8, device=t.device)
    N.check(e.lib.vcp_base64(e.handle, t.data_ptr() if n else None, n, out.data_ptr() if n else None))
    return out[:4 * ((n + 2) // 3)].cpu().numpy().tobytes()
