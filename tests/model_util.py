"""Loader for the sequential CPU model of the GPU deflate (tests/model/deflate_model.c). Test infrastructure."""
import ctypes as C
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = os.path.join(HERE, "model", "deflate_model.c")
SO = os.path.join(HERE, "model", "libdeflate_model.so")


class Params(C.Structure):
    _fields_ = [(n, C.c_int32) for n in ("bpp", "hash_bits", "ways", "lane_cap", "too_far", "lazy", "cont_min",
                                         "prime_bytes", "capped_wins", "inwin", "cont_maxd", "sub_bytes", "hash2_bytes",
                                         "hash2_bits", "noisy_thresh", "noisy_minlen", "noisy_neard", "cost_maxlen",
                                         "cost_margin", "cost_warm", "hash2_ways", "ins_limit", "noisy_ways1", "lane_cap_win", "rowlen", "row_gate", "exact_sel", "group_subs", "prime_win", "cost_epoch")] + [("block_bytes", C.c_int64)]


class Stats(C.Structure):
    _fields_ = [("tokens", C.c_int64), ("blocks", C.c_int64), ("stored_blocks", C.c_int64)]


# the configuration csrc/deflate_lz.cu + deflate_huff.cu implement (group_subs mirrors kGroupSubs in vcp_internal.cuh)
KERNEL_PARAMS = dict(bpp=3, hash_bits=10, ways=2, lane_cap=16, too_far=32768, lazy=16, cont_min=258, prime_bytes=16384,
                     capped_wins=1, inwin=0, cont_maxd=1, noisy_thresh=160, noisy_minlen=0, noisy_neard=0, cost_maxlen=8,
                     cost_margin=0, cost_warm=64, hash2_ways=2, ins_limit=0, noisy_ways1=1, lane_cap_win=0, rowlen=-1, row_gate=1, exact_sel=1, group_subs=1, prime_win=512, cost_epoch=64, hash2_bytes=4, hash2_bits=10, sub_bytes=16384,
                     block_bytes=512 * 1024)


# effort classes of csrc/deflate_lz.cu (CfgFast / CfgDefault / CfgBest), selected by compress_level
LEVEL_PARAMS = {
    "fast": dict(hash_bits=9, hash2_bits=9, noisy_thresh=-1, lazy=0, rowlen=0),
    "default": {},
    "best": dict(hash_bits=11, hash2_bits=12, noisy_thresh=0, row_gate=0),
}


def params_for_level(level: int) -> dict:
    return dict(LEVEL_PARAMS["best" if level >= 7 else "fast" if 1 <= level <= 3 else "default"])


def load():
    if not os.path.exists(SO) or os.path.getmtime(SO) < os.path.getmtime(SRC):
        subprocess.run(["gcc", "-O2", "-shared", "-fPIC", "-o", SO, SRC], check=True)
    lib = C.CDLL(SO)
    lib.dm_deflate_page.restype = C.c_int64
    lib.dm_deflate_page.argtypes = [C.c_void_p, C.c_int64, C.POINTER(Params), C.c_void_p, C.c_void_p, C.POINTER(Stats)]
    lib.dm_lz_subchunk_ex.restype = C.c_int64
    lib.dm_lz_subchunk_ex.argtypes = [C.c_void_p, C.c_int64, C.c_int64, C.c_int64, C.POINTER(Params), C.c_void_p, C.c_void_p,
                                      C.c_void_p, C.c_void_p, C.c_int]
    return lib


def _params(kw):
    p = dict(KERNEL_PARAMS); p.update(kw)
    if p["rowlen"] < 0:                  # a bare stream goes through the kernels as a 1 x 1 "page" of bpp channels
        p["rowlen"] = 1 + p["bpp"]
    return p


def deflate(lib, stream: bytes, **kw):
    p = _params(kw)
    P = Params(**p)
    src = np.concatenate([np.frombuffer(stream, np.uint8), np.zeros(512, np.uint8)])
    out = np.zeros(len(stream) + len(stream) // 8 + 4096, np.uint8)
    st = Stats()
    n = lib.dm_deflate_page(src.ctypes.data, len(stream), C.byref(P), out.ctypes.data, None, C.byref(st))
    return out[:n].tobytes(), st


def lz_tokens(lib, stream: bytes, **kw):
    """Per sub-chunk token lists as the kernel lays them out: list of (tokens uint32 array, hist[316])."""
    p = _params(kw)
    P = Params(**p)
    F = len(stream)
    src = np.concatenate([np.frombuffer(stream, np.uint8), np.zeros(512, np.uint8)])
    res = []
    T = np.zeros(p["ways"] << p["hash_bits"], np.uint16)
    T2 = np.zeros((max(1, p["hash2_ways"]) << p["hash2_bits"]) + 1, np.uint16)
    G = max(1, p["group_subs"])
    idx = 0
    for bs in range(0, F, p["block_bytes"]):
        be = min(F, bs + p["block_bytes"])
        for s in range(bs, be, p["sub_bytes"]):
            e = min(be, s + p["sub_bytes"])
            tok = np.zeros(e - s + 64, np.uint32); hist = np.zeros(316, np.uint32)
            n = lib.dm_lz_subchunk_ex(src.ctypes.data, F, s, e, C.byref(P), tok.ctypes.data, hist.ctypes.data,
                                      T.ctypes.data, T2.ctypes.data, int(idx % G != 0))
            res.append((tok[:n].copy(), hist))
            idx += 1
    return res
