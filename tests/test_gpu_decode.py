"""GPU: PNG decode (inflate + un-filter) against Pillow, for this library's PNGs and for Pillow's own."""
import io
import zlib

import numpy as np
import pytest
from PIL import Image

from tests import util as U

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def V():
    import vision_compression_project_b200 as v
    return v


def _px(im):
    a = np.asarray(im)
    return a.reshape(im.height, im.width, -1)


def _pillow(png):
    """What Image.open(png).load() does with these bytes: its pixels, or None when it raises."""
    try:
        im = Image.open(io.BytesIO(png))
        im.load()
        return _px(im).copy()
    except Exception:
        return None


def _assert_like_pillow(V, pngs, names=None):
    """The decode contract: a ValueError exactly where Pillow raises, Pillow's pixels everywhere else.  Returns the number of rejects."""
    dec = V.decode_pages(pngs)
    n_err = 0
    for k, (png, d) in enumerate(zip(pngs, dec)):
        want = _pillow(png)
        tag = names[k] if names else k
        if want is None:
            assert isinstance(d, ValueError), f"{tag}: Pillow rejects this file, the GPU decoder returned pixels"
            n_err += 1
        else:
            assert not isinstance(d, Exception), f"{tag}: Pillow decodes this file, the GPU decoder rejected it ({d})"
            assert d.shape == want.shape and np.array_equal(d, want), f"{tag}: pixels differ from Pillow's"
    return n_err


def test_decode_own_pngs_round_trip(V, ref_page):
    from vision_compression_project_b200 import synth
    pages = [ref_page, synth.make_page(1, "letter", 200, photo=True), synth.make_page(2, size=(333, 517), mode="L"),
             Image.fromarray(np.random.default_rng(0).integers(0, 256, (300, 400, 3), dtype=np.uint8), "RGB")]   # incl. stored blocks
    res = V.prepare_pages(pages, mode=None, want_base64=False)
    dec = V.decode_pages([r.png for r in res])
    for im, d in zip(pages, dec):
        assert not isinstance(d, Exception)
        assert np.array_equal(d, _px(im))
    dev = V.decode_pages([res[0].png], to_device=True)[0]
    assert dev.is_cuda and np.array_equal(dev.cpu().numpy(), _px(ref_page))


def test_decode_pillow_pngs_all_filters_and_modes(V):
    rng = np.random.default_rng(3)
    pngs, exp = [], []
    for t in range(40):
        h, w = int(rng.integers(1, 90)), int(rng.integers(1, 90))
        mode = ["L", "RGB", "RGBA", "LA"][t % 4]
        c = {"L": 1, "LA": 2, "RGB": 3, "RGBA": 4}[mode]
        kind = t % 3
        if kind == 0:
            px = rng.integers(0, 256, (h, w, c), dtype=np.uint8)
        elif kind == 1:
            px = np.full((h, w, c), int(rng.integers(0, 256)), np.uint8)
        else:
            px = (np.add.outer(np.arange(h) * 3, np.arange(w) * 5)[:, :, None] + np.arange(c) * 17).astype(np.uint8)
        im = Image.fromarray(px[:, :, 0] if c == 1 else px, mode)
        kw = [{}, {"optimize": True}, {"compress_level": 0}, {"compress_level": 1}, {"compress_level": 9}][t % 5]
        pngs.append(U.pillow_png(im, **kw)); exp.append(px)
    # hand-made streams: fixed-Huffman blocks (zlib picks them for tiny inputs) and every filter type in one image
    px = rng.integers(0, 256, (5, 7, 3), dtype=np.uint8)
    from oracle import restate as R
    rows = []
    for y in range(5):
        prev = px[y - 1].reshape(-1) if y else np.zeros(21, np.uint8)
        cur = px[y].reshape(-1).astype(np.int32)
        left = np.concatenate([np.zeros(3, np.int32), cur[:-3]]); ul = np.concatenate([np.zeros(3, np.int32), prev[:-3].astype(np.int32)])
        pred = [np.zeros(21, np.int32), left, prev.astype(np.int32), (left + prev) >> 1, R._paeth(left, prev, ul)][y]
        rows.append(bytes([y]) + ((cur - pred) & 255).astype(np.uint8).tobytes())
    pngs.append(R.png_wrap(7, 5, 3, zlib.compress(b"".join(rows), 9))); exp.append(px)
    dec = V.decode_pages(pngs)
    for k, (d, e) in enumerate(zip(dec, exp)):
        assert not isinstance(d, Exception), k
        assert np.array_equal(d, e), k


def test_decode_recorded_reference_png_and_errors(V, golden_dir, ref_page):
    import os
    raw = open(os.path.join(golden_dir, "ref_page_1.png"), "rb").read()        # written by another zlib build (the reference's run)
    bad_crc_ok = bytearray(raw)
    trunc = raw[:len(raw) // 2]
    corrupt = bytearray(raw); corrupt[5000:5040] = bytes(40)
    dec = V.decode_pages([raw, b"not a png", trunc, bytes(corrupt)])
    assert np.array_equal(dec[0], _px(ref_page))
    assert all(isinstance(d, ValueError) for d in dec[1:4])
    assert _assert_like_pillow(V, [raw, b"not a png", trunc, bytes(corrupt)]) == 3
    pal = io.BytesIO(); Image.new("P", (4, 4)).save(pal, format="PNG")
    assert isinstance(V.decode_pages([pal.getvalue()])[0], ValueError)


def _png_from_idats(w, h, c, idats):
    from oracle import restate as R
    import struct
    out = [R.PNG_SIG, R.png_chunk(b"IHDR", struct.pack(">IIBBBBB", w, h, 8, R.COLOR_TYPE[c], 0, 0, 0))]
    out += [R.png_chunk(b"IDAT", d) for d in idats]
    out.append(R.png_chunk(b"IEND", b""))
    return b"".join(out)


def _filtered(px):
    """Filter type 4/2/1/3/0 cycling per row (restated Paeth etc. from the oracle), returns the filtered stream."""
    from oracle import restate as R
    h, w, c = px.shape
    nb = w * c
    rows = []
    for y in range(h):
        prev = px[y - 1].reshape(-1).astype(np.int32) if y else np.zeros(nb, np.int32)
        cur = px[y].reshape(-1).astype(np.int32)
        left = np.concatenate([np.zeros(c, np.int32), cur[:-c]]); ul = np.concatenate([np.zeros(c, np.int32), prev[:-c]])
        ft = (4, 2, 1, 3, 0)[y % 5]
        pred = {0: np.zeros(nb, np.int32), 1: left, 2: prev, 3: (left + prev) >> 1, 4: R._paeth(left, prev, ul)}[ft]
        rows.append(bytes([ft]) + ((cur - pred) & 255).astype(np.uint8).tobytes())
    return b"".join(rows)


def test_decode_segment_parallel_foreign_streams(V):
    """IDATs that each hold whole deflate blocks (zlib Z_SYNC_FLUSH cuts, history kept across the cut) take the
    segment-parallel path; segments shorter than the 32 KiB window chain their windows; cuts that are not block
    boundaries, or an Adler-32 in its own IDAT, must fall back to the serial path with the same pixels."""
    rng = np.random.default_rng(11)
    pngs, exp = [], []
    for t, (h, w, c, piece) in enumerate([(240, 301, 3, 5000), (64, 4000, 1, 70000), (500, 257, 4, 1000), (130, 97, 2, 40000), (90, 1000, 3, 33000)]):
        base = rng.integers(0, 256, (8, w, c), dtype=np.uint8)
        px = np.concatenate([base[rng.permutation(8)] for _ in range((h + 7) // 8)])[:h].copy()      # rows repeat: long far matches
        px[rng.integers(0, h, 40), rng.integers(0, w, 40)] ^= 0x55
        filt = _filtered(px)
        co = zlib.compressobj(6)
        idats = []
        for off in range(0, len(filt), piece):
            idats.append(co.compress(filt[off:off + piece]) + co.flush(zlib.Z_SYNC_FLUSH))
        tail = co.flush()
        variant = t % 3
        if variant == 0:
            idats[-1] += tail                                   # final (empty) block + Adler in the last IDAT: fully parallel
        elif variant == 1:
            idats.append(tail)                                  # final block + Adler as their own IDAT: still whole blocks
        else:
            idats[-1] += tail[:-2]; idats.append(tail[-2:])     # Adler split across IDATs: not parallel, serial fallback
        assert zlib.decompress(b"".join(idats)) == filt
        pngs.append(_png_from_idats(w, h, c, idats)); exp.append(px)
    # cuts in the middle of blocks
    px = rng.integers(0, 4, (200, 300, 3), dtype=np.uint8) * 60
    z = zlib.compress(_filtered(px), 6)
    pngs.append(_png_from_idats(300, 200, 3, [z[i:i + 777] for i in range(0, len(z), 777)])); exp.append(px)
    dec = V.decode_pages(pngs)
    for k, (d, e) in enumerate(zip(dec, exp)):
        assert not isinstance(d, Exception), k
        assert np.array_equal(d, e), k


def test_decode_large_batch_mixed_and_corrupt_segment(V):
    from vision_compression_project_b200 import synth
    pages = [synth.make_page(i, "letter", 200, photo=(i % 2 == 1)) for i in range(4)]
    pages.append(synth.make_page(5, size=(2000, 2600), mode="L"))
    pages.append(Image.fromarray(np.random.default_rng(5).integers(0, 256, (1200, 900, 4), dtype=np.uint8), "RGBA"))
    ours = [r.png for r in V.prepare_pages(pages, mode=None, want_base64=False)]
    pil = [U.pillow_png(p) for p in pages]
    bad = bytearray(ours[0]); bad[len(bad) // 2:len(bad) // 2 + 64] = bytes(64)      # inside one IDAT of a multi-IDAT PNG
    batch = ours + pil + [bytes(bad)] + ours
    dec = V.decode_pages(batch, to_device=True)
    for k, d in enumerate(dec):
        if k == 2 * len(pages):
            assert isinstance(d, ValueError) and _pillow(batch[k]) is None        # rejected, as Pillow does
            continue
        assert not isinstance(d, Exception), k
        assert np.array_equal(d.cpu().numpy(), _px(pages[k % len(pages) if k < 2 * len(pages) else k - 2 * len(pages) - 1])), k


def test_decode_two_groups_keep_page_order(V):
    """>= 128 decodable pages run as two page groups on two streams: results stay in input order, a bad page stays in its slot."""
    rng = np.random.default_rng(21)
    ims = []
    for t in range(150):
        h, w = int(rng.integers(20, 120)), int(rng.integers(20, 160))
        mode = ["L", "RGB", "RGBA"][t % 3]
        c = {"L": 1, "RGB": 3, "RGBA": 4}[mode]
        px = (rng.integers(0, 3, (h, w, c)) * 90 + t).astype(np.uint8)
        ims.append(Image.fromarray(px[:, :, 0] if c == 1 else px, mode))
    ours = [r.png for r in V.prepare_pages(ims[:75], mode=None, want_base64=False)]
    pngs = ours + [U.pillow_png(im) for im in ims[75:]]
    pngs.insert(40, b"\x89PNG\r\n\x1a\n" + bytes(40))
    dec = V.decode_pages(pngs)
    assert isinstance(dec[40], ValueError)
    del dec[40]
    for k, (d, im) in enumerate(zip(dec, ims)):
        assert not isinstance(d, Exception), k
        assert np.array_equal(d, _px(im)), k


@pytest.mark.timeout(300)
def test_decode_corrupted_streams_like_pillow(V):
    """Random damage anywhere in the file (container, zlib header, Huffman tables, tokens, Adler-32, chunk lengths and CRCs): every PNG of
    the batch is rejected exactly when Image.open(png).load() raises, and otherwise decodes to Pillow's pixels; intact neighbours stay
    exact; nothing hangs."""
    from vision_compression_project_b200 import synth
    rng = np.random.default_rng(99)
    page = synth.make_page(7, size=(700, 900), photo=True)
    good = [V.prepare_pages([page], want_base64=False)[0].png, U.pillow_png(page)]
    batch, names = [], []
    for t in range(96):
        b = bytearray(good[t % 2])
        mode = t % 6
        if mode == 0:
            for _ in range(3):
                b[int(rng.integers(33, len(b)))] ^= int(rng.integers(1, 256))           # behind IHDR: the geometry stays
        elif mode == 1:
            o = int(rng.integers(60, len(b) - 200)); b[o:o + 64] = rng.integers(0, 256, 64, dtype=np.uint8).tobytes()
        elif mode == 2:
            del b[int(rng.integers(100, len(b) - 100)):]
        elif mode == 3:
            o = int(rng.integers(41, 120)); b[o] ^= 0xFF                      # inside the first block header
        elif mode == 4:
            o = len(b) - int(rng.integers(13, 40)); b[o] ^= 1 << int(rng.integers(0, 8))   # the last tokens / end-of-block / Adler-32 / IEND
        else:
            o = int(rng.integers(8, 33)); b[o] ^= 1 << int(rng.integers(0, 8))             # IHDR: length, tag, fields, CRC
        batch.append(bytes(b)); names.append(f"case {t} (mode {mode})")
        if t % 8 == 7:
            batch.append(good[(t // 8) % 2]); names.append("intact")
    n_err = _assert_like_pillow(V, batch, names)
    assert n_err >= 48, n_err                                                  # the damage is real: most of it must be detected


def test_decode_accepts_and_rejects_what_pillow_does(V):
    """Hand-made files for every rule of Pillow's PNG loader that is not 'the stream is valid': chunk CRCs are checked in front of
    the image data only; the Adler-32 is checked only when inflate() meets it in the call that produced the last row; data behind
    the last row are ignored; a stream that ends early leaves black rows; zlib's header and code-completeness rules."""
    import struct
    from oracle import restate as R

    def chunk(tag, data, crc=None):
        c = zlib.crc32(tag + data) if crc is None else crc
        return struct.pack(">I", len(data)) + tag + data + struct.pack(">I", c & 0xFFFFFFFF)

    def png(w, h, c, idats, ihdr_crc=None, idat_crc=None, iend=True, pre=b"", post=b""):
        out = [R.PNG_SIG, chunk(b"IHDR", struct.pack(">IIBBBBB", w, h, 8, R.COLOR_TYPE[c], 0, 0, 0), ihdr_crc), pre]
        out += [chunk(b"IDAT", d, idat_crc) for d in idats]
        out.append(post)
        if iend:
            out.append(chunk(b"IEND", b""))
        return b"".join(out)

    rng = np.random.default_rng(1)
    cases = {}
    for tag, (w, h, c) in {"small": (30, 20, 3), "rows": (700, 300, 3), "gray": (1000, 150, 1)}.items():
        px = (rng.integers(0, 6, (h, w, c)) * 40).astype(np.uint8)
        filt = _filtered(px)
        rowlen = 1 + w * c
        z = zlib.compress(filt, 6)
        mk = lambda idats, **kw: png(w, h, c, idats, **kw)
        cases[f"{tag}: good"] = mk([z])
        cases[f"{tag}: bad IDAT crc (not checked)"] = mk([z], idat_crc=123)
        cases[f"{tag}: bad IHDR crc"] = mk([z], ihdr_crc=123)
        bad = bytearray(z); bad[-1] ^= 1
        cases[f"{tag}: bad adler"] = mk([bytes(bad)])
        cases[f"{tag}: bad adler in its own IDAT (never read)"] = mk([bytes(bad[:-4]), bytes(bad[-4:])])
        cases[f"{tag}: adler split over two IDATs, wrong"] = mk([bytes(bad[:-2]), bytes(bad[-2:])])
        cases[f"{tag}: no IEND"] = mk([z], iend=False)
        cases[f"{tag}: adler cut"] = mk([z[:-1]])
        cases[f"{tag}: adler missing"] = mk([z[:-4]])
        cases[f"{tag}: garbage behind the stream"] = mk([z + b"abcdef"])
        cases[f"{tag}: garbage IDAT behind the stream"] = mk([z, b"zzzzzzzz"])
        cases[f"{tag}: one extra row"] = mk([zlib.compress(filt + bytes(rowlen), 6)])
        cases[f"{tag}: extra partial row"] = mk([zlib.compress(filt + bytes(rowlen // 3), 6)])
        cases[f"{tag}: one row short"] = mk([zlib.compress(filt[:-rowlen], 6)])
        cases[f"{tag}: half the rows"] = mk([zlib.compress(filt[:rowlen * (h // 2) + 5], 6)])
        short_bad = bytearray(zlib.compress(filt[:-rowlen], 6)); short_bad[-2] ^= 4
        cases[f"{tag}: one row short, bad adler"] = mk([bytes(short_bad)])
        cases[f"{tag}: one row short, adler missing"] = mk([zlib.compress(filt[:-rowlen], 6)[:-4]])
        f5 = bytearray(filt); f5[rowlen * 3] = 5
        cases[f"{tag}: filter type 5"] = mk([zlib.compress(bytes(f5), 6)])
        for cinfo in (8, 6):
            zz = bytearray(z); zz[0] = (cinfo << 4) | 8
            zz[1] = next(f for f in range(256) if ((zz[0] << 8) | f) % 31 == 0 and not f & 0x20)
            cases[f"{tag}: window bits {cinfo + 8}"] = mk([bytes(zz)])
        zz = bytearray(z); zz[1] ^= 1
        cases[f"{tag}: bad zlib header check"] = mk([bytes(zz)])
        cases[f"{tag}: unknown chunk in front, good crc"] = mk([z], pre=chunk(b"abCd", b"hello"))
        cases[f"{tag}: unknown chunk in front, bad crc"] = mk([z], pre=chunk(b"abCd", b"hello", 5))
        cases[f"{tag}: text chunk behind, bad crc"] = mk([z], post=chunk(b"tEXt", b"k\0v", 5))
        cases[f"{tag}: junk behind the data"] = mk([z], post=b"\0\0\0\1\xff\xff\xff\xffAAAAA")
        cases[f"{tag}: empty IDAT in the middle"] = mk([z[:10], b"", z[10:]])
        cases[f"{tag}: other chunk between IDATs (data end there)"] = png(w, h, c, [z[:len(z) // 2]], post=chunk(b"tEXt", b"k\0v") + chunk(b"IDAT", z[len(z) // 2:]))
        co = zlib.compressobj(0)
        cases[f"{tag}: stored blocks"] = mk([co.compress(filt) + co.flush()])
        g = mk([z])
        cases[f"{tag}: truncated in the data"] = g[:len(g) - 40]
        cases[f"{tag}: truncated in IEND"] = g[:len(g) - 6]
        cases[f"{tag}: IDAT longer than the file"] = g[:33] + struct.pack(">I", len(z) + 1000) + g[37:]
        co = zlib.compressobj(6)
        body = co.compress(filt) + co.flush(zlib.Z_SYNC_FLUSH)
        tail = co.flush()                                                   # empty final block + Adler-32
        cases[f"{tag}: empty final block, good"] = mk([body + tail])
        t2 = bytearray(tail); t2[-1] ^= 0x10
        cases[f"{tag}: empty final block, bad adler"] = mk([body + bytes(t2)])
        cases[f"{tag}: empty final block in its own IDAT, bad adler"] = mk([body, bytes(t2)])
        cases[f"{tag}: bad block type behind the last row"] = mk([body + b"\x07\x00\x00"])
    names = list(cases)
    n_err = _assert_like_pillow(V, [cases[k] for k in names], names)
    assert n_err >= 3 * 9


def test_decode_png_cut_into_thousands_of_idats(V):
    """IDAT starts are optional parse units: a stream cut into 5000 tiny IDATs decodes like any other."""
    rng = np.random.default_rng(5)
    px = (rng.integers(0, 5, (300, 400, 3)) * 50).astype(np.uint8)
    z = zlib.compress(_filtered(px), 6)
    step = max(1, len(z) // 5000)
    png = _png_from_idats(400, 300, 3, [z[i:i + step] for i in range(0, len(z), step)])
    d = V.decode_pages([png])[0]
    assert not isinstance(d, Exception) and np.array_equal(d, px)


def _filtered_png(px, types):
    """A PNG of `px` (H, W, C) whose row y uses filter types[y % len(types)] (hand-filtered, then zlib)."""
    import struct
    h, w, c = px.shape
    rows = []
    prev = np.zeros(w * c, np.int32)
    for y in range(h):
        cur = px[y].reshape(-1).astype(np.int32)
        a = np.concatenate([np.zeros(c, np.int32), cur[:-c]])
        cc = np.concatenate([np.zeros(c, np.int32), prev[:-c]])
        t = types[y % len(types)]
        if t == 0: pred = 0
        elif t == 1: pred = a
        elif t == 2: pred = prev
        elif t == 3: pred = (a + prev) >> 1
        else:
            p = a + prev - cc
            pa, pb, pc = np.abs(p - a), np.abs(p - prev), np.abs(p - cc)
            pred = np.where((pa <= pb) & (pa <= pc), a, np.where(pb <= pc, prev, cc))
        rows.append(bytes([t]) + ((cur - pred) & 255).astype(np.uint8).tobytes())
        prev = cur
    def chunk(t, d): return struct.pack(">I", len(d)) + t + d + struct.pack(">I", zlib.crc32(t + d))
    ct = {1: 0, 2: 4, 3: 2, 4: 6}[c]
    return (b"\x89PNG\r\n\x1a\n" + chunk(b"IHDR", struct.pack(">IIBBBBB", w, h, 8, ct, 0, 0, 0)) +
            chunk(b"IDAT", zlib.compress(b"".join(rows), 6)) + chunk(b"IEND", b""))


def test_unfilter_geometries_and_filter_mixes(V):
    """The un-filter kernel's edges: widths around its 32-pixel chunks and 3-pixel word rule, heights around its 32-row bands and
    4-band groups, all four channel counts, every filter type alone and mixed row by row — against Pillow's decode of the same bytes."""
    rng = np.random.default_rng(11)
    widths = [1, 2, 3, 4, 5, 7, 29, 31, 32, 33, 34, 35, 36, 61, 63, 64, 65, 67, 95, 97, 129, 257, 400, 641]
    heights = [1, 2, 31, 32, 33, 63, 64, 65, 96, 127, 128, 129, 160, 200, 257]
    mixes = [[0], [1], [2], [3], [4], [4, 3], [2, 2, 2, 4], [0, 1, 2, 3, 4], [1, 2], [3, 3, 1, 4, 2, 0, 4]]
    pngs, names = [], []
    k = 0
    for w in widths:
        for c in (1, 2, 3, 4):
            h = heights[k % len(heights)]; mix = mixes[k % len(mixes)]; k += 1
            if k % 3 == 0:
                px = rng.integers(0, 256, (h, w, c), dtype=np.uint8)
            else:
                base = (np.add.outer(np.arange(h) * 5, np.arange(w) * 3)[:, :, None] + np.arange(c) * 29)
                px = (base + rng.integers(0, 4, (h, w, c))).astype(np.uint8)
            pngs.append(_filtered_png(px, mix)); names.append(f"{w}x{h}x{c} filters {mix}")
    assert _assert_like_pillow(V, pngs, names) == 0
    # one tall page per channel count: many groups of bands, row-by-row filter mix, width not a multiple of anything
    pngs, names = [], []
    for c in (1, 2, 3, 4):
        px = rng.integers(0, 256, (700, 1003, c), dtype=np.uint8)
        px[100:300] = 255                                                  # blank rows: bands without Avg / Paeth rows use the packed step
        pngs.append(_filtered_png(px, [2] * 40 + [4, 1, 3, 0, 2, 4, 4, 3])); names.append(f"tall {c}")
    assert _assert_like_pillow(V, pngs, names) == 0
