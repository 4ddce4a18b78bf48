"""CPU: the rule the GPU decoder uses to find deflate block starts (png_decode.cu k_infl_scan1/2), restated in oracle/inflate_walk.py,
against streams written by zlib (what Pillow's ZipEncode.c drives): every dynamic block zlib starts is accepted, and nothing else is."""
import zlib

import numpy as np

from oracle import inflate_walk as W


def _streams():
    rng = np.random.default_rng(4)
    text = bytes(rng.integers(97, 123, 60000, dtype=np.uint8))
    sparse = np.zeros(90000, np.uint8); sparse[rng.integers(0, 90000, 3000)] = rng.integers(1, 255, 3000); sparse = sparse.tobytes()
    yield "text-6", zlib.compress(text, 6), len(text)
    yield "sparse-9", zlib.compress(sparse, 9), len(sparse)
    co = zlib.compressobj(6, zlib.DEFLATED, 15, 9, zlib.Z_FILTERED)          # Pillow's configuration
    pieces = [co.compress(text[:30000]), co.flush(zlib.Z_SYNC_FLUSH), co.compress(sparse[:40000]), co.flush(zlib.Z_FULL_FLUSH),
              co.compress(text[30000:]), co.flush()]
    yield "flushed", b"".join(pieces), 30000 + 40000 + 30000
    co = zlib.compressobj(1, zlib.DEFLATED, 15, 1)                           # memLevel 1: a new block every 128 symbols... many blocks
    yield "many-blocks", co.compress(text[:20000]) + co.flush(), 20000


def test_walk_matches_zlib_and_every_dynamic_block_is_recognised():
    for name, z, n in _streams():
        blocks, out_len = W.block_starts(z)
        assert out_len == n == len(zlib.decompress(z)), name
        dyn = [b for b, t, _ in blocks if t == 2]
        assert dyn, name
        for b in dyn:
            if b >= 17:
                assert W.looks_like_dynamic_header(z, b), (name, b)
        outs = [o for _, _, o in blocks]
        assert outs == sorted(outs), name


def test_no_false_block_starts_in_real_streams():
    checked = 0
    for name, z, _ in _streams():
        blocks, _ = W.block_starts(z)
        true = {b for b, _, _ in blocks}
        nbits = 8 * len(z)
        for bit in range(17, nbits, 7 if nbits > 100000 else 3):            # a stride keeps the pure-Python walk short
            if bit in true:
                continue
            checked += 1
            assert not W.looks_like_dynamic_header(z, bit), (name, bit)
    assert checked > 50000
