"""TEST INFRASTRUCTURE ONLY (see oracle/__init__.py): a plain-Python walk over a zlib / deflate stream (RFC 1950 / 1951; the
algorithm zlib's inflate.c runs for ZipDecode.c) that records where every deflate block starts, and a restatement of the rule
png_decode.cu uses to recognise a dynamic-Huffman block header at an arbitrary bit position (k_infl_scan1 / k_infl_scan2).
Pure-Python loops: small streams only.  Nothing in the product imports this."""
from __future__ import annotations

CL_ORDER = (16, 17, 18, 0, 8, 7, 9, 6, 10, 5, 11, 4, 12, 3, 13, 2, 14, 1, 15)
LEN_BASE = (3, 4, 5, 6, 7, 8, 9, 10, 11, 13, 15, 17, 19, 23, 27, 31, 35, 43, 51, 59, 67, 83, 99, 115, 131, 163, 195, 227, 258)
LEN_EXTRA = (0, 0, 0, 0, 0, 0, 0, 0, 1, 1, 1, 1, 2, 2, 2, 2, 3, 3, 3, 3, 4, 4, 4, 4, 5, 5, 5, 5, 0)
DIST_EXTRA = (0, 0, 0, 0, 1, 1, 2, 2, 3, 3, 4, 4, 5, 5, 6, 6, 7, 7, 8, 8, 9, 9, 10, 10, 11, 11, 12, 12, 13, 13)


class _Bits:
    def __init__(self, data: bytes, bit: int = 0):
        self.d, self.p = data, bit

    def get(self, n: int) -> int:
        v = 0
        for i in range(n):
            byte = self.d[(self.p + i) >> 3] if ((self.p + i) >> 3) < len(self.d) else 0
            v |= ((byte >> ((self.p + i) & 7)) & 1) << i
        self.p += n
        return v


def _canonical(lengths):
    """code lengths -> {(length, code): symbol} (codes MSB-first as RFC 1951 assigns them)"""
    count = [0] * 16
    for l in lengths:
        count[l] += 1
    count[0] = 0
    code, nxt = 0, [0] * 16
    for l in range(1, 16):
        code = (code + count[l - 1]) << 1
        nxt[l] = code
    table = {}
    for s, l in enumerate(lengths):
        if l:
            table[(l, nxt[l])] = s
            nxt[l] += 1
    return table


def _decode(bits: _Bits, table) -> int:
    code = 0
    for l in range(1, 16):
        code = (code << 1) | bits.get(1)
        if (l, code) in table:
            return table[(l, code)]
    raise ValueError("bad code")


def _read_lengths(bits: _Bits):
    nll, nd, ncl = bits.get(5) + 257, bits.get(5) + 1, bits.get(4) + 4
    cl = [0] * 19
    for i in range(ncl):
        cl[CL_ORDER[i]] = bits.get(3)
    t = _canonical(cl)
    lens = []
    while len(lens) < nll + nd:
        s = _decode(bits, t)
        if s < 16:
            lens.append(s)
        elif s == 16:
            if not lens:
                raise ValueError("repeat with nothing to repeat")
            lens += [lens[-1]] * (3 + bits.get(2))
        elif s == 17:
            lens += [0] * (3 + bits.get(3))
        else:
            lens += [0] * (11 + bits.get(7))
    if len(lens) > nll + nd:
        raise ValueError("run past the end")
    return lens[:nll], lens[nll:]


def block_starts(zstream: bytes):
    """[(bit position, btype, output position)] of every deflate block of a zlib stream, and the inflated length."""
    bits = _Bits(zstream, 16)
    out_len, blocks = 0, []
    while True:
        start = bits.p
        last, btype = bits.get(1), bits.get(2)
        blocks.append((start, btype, out_len))
        if btype == 0:
            bits.p = (bits.p + 7) & ~7
            ln = bits.get(16); bits.get(16)
            bits.p += 8 * ln; out_len += ln
        else:
            if btype == 1:
                ll = [8] * 144 + [9] * 112 + [7] * 24 + [8] * 8; dd = [5] * 30
            else:
                ll, dd = _read_lengths(bits)
            tl, td = _canonical(ll), _canonical(dd)
            while True:
                s = _decode(bits, tl)
                if s < 256:
                    out_len += 1
                elif s == 256:
                    break
                else:
                    out_len += LEN_BASE[s - 257] + bits.get(LEN_EXTRA[s - 257])
                    bits.get(DIST_EXTRA[_decode(bits, td)])
        if last:
            return blocks, out_len


def looks_like_dynamic_header(zstream: bytes, bit: int) -> bool:
    """The acceptance rule of k_infl_scan1 + k_infl_scan2 at one bit position."""
    nbits = 8 * len(zstream)
    if bit < 17 or bit + 60 > nbits:
        return False
    b = _Bits(zstream, bit)
    b.get(1)
    if b.get(2) != 2:
        return False
    hlit, hdist, ncl = b.get(5), b.get(5), b.get(4) + 4
    if hlit > 29 or hdist > 29:
        return False
    cl = [0] * 19
    for i in range(ncl):
        cl[CL_ORDER[i]] = b.get(3)
    if sum(128 >> l for l in cl if l) != 128:                       # the code-length code must be complete
        return False
    t = _canonical(cl)
    nll, nd = hlit + 257, hdist + 1
    kr_ll = kr_d = n_d = 0
    i, prev, has_eob = 0, 0, False
    while i < nll + nd:
        if b.p + 14 > nbits:
            return False
        s = _decode(b, t)
        if s < 16:
            rep, v = 1, s
        elif s == 16:
            if i == 0:
                return False
            rep, v = 3 + b.get(2), prev
        elif s == 17:
            rep, v = 3 + b.get(3), 0
        else:
            rep, v = 11 + b.get(7), 0
        if i + rep > nll + nd:
            return False
        if v:
            for r in range(rep):
                if i + r < nll:
                    kr_ll += 32768 >> v
                    has_eob |= i + r == 256
                else:
                    kr_d += 32768 >> v; n_d += 1
            if kr_ll > 32768 or kr_d > 32768:
                return False
        prev = v; i += rep
    return has_eob and kr_ll == 32768 and (kr_d == 32768 or n_d <= 1)
