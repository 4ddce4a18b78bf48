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
    assert all(isinstance(d, ValueError) for d in dec[1:3])
    assert isinstance(dec[3], ValueError) or not np.array_equal(dec[3], _px(ref_page))
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
            assert isinstance(d, ValueError) or not np.array_equal(d.cpu().numpy(), _px(pages[0]))
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


@pytest.mark.timeout(180)
def test_decode_corrupted_streams_end_in_a_status(V):
    """Random damage anywhere in the file (container, zlib header, Huffman tables, tokens, Adler): every PNG of the batch comes back
    as an error or as some array of the right shape, the intact neighbours stay exact, nothing hangs."""
    from vision_compression_project_b200 import synth
    rng = np.random.default_rng(99)
    page = synth.make_page(7, size=(700, 900), photo=True)
    good = [V.prepare_pages([page], want_base64=False)[0].png, U.pillow_png(page)]
    batch, kind = [], []
    for t in range(48):
        b = bytearray(good[t % 2])
        mode = t % 4
        if mode == 0:
            for _ in range(3):
                b[int(rng.integers(33, len(b)))] ^= int(rng.integers(1, 256))           # behind IHDR: the geometry stays
        elif mode == 1:
            o = int(rng.integers(60, len(b) - 200)); b[o:o + 64] = rng.integers(0, 256, 64, dtype=np.uint8).tobytes()
        elif mode == 2:
            del b[int(rng.integers(100, len(b) - 100)):]
        else:
            o = int(rng.integers(41, 120)); b[o] ^= 0xFF                      # inside the first block header
        batch.append(bytes(b)); kind.append(mode)
        if t % 8 == 7:
            batch.append(good[(t // 8) % 2]); kind.append(-1)
    dec = V.decode_pages(batch)
    exp = _px(page)
    n_err = 0
    for d, k in zip(dec, kind):
        if k == -1:
            assert not isinstance(d, Exception) and np.array_equal(d, exp)
        elif isinstance(d, Exception):
            n_err += 1
        else:
            assert d.shape == exp.shape
    assert n_err >= 24                                                         # most damage is detected (chunk CRCs are not checked)


def test_decode_png_cut_into_thousands_of_idats(V):
    """IDAT starts are optional parse units: a stream cut into 5000 tiny IDATs decodes like any other."""
    rng = np.random.default_rng(5)
    px = (rng.integers(0, 5, (300, 400, 3)) * 50).astype(np.uint8)
    z = zlib.compress(_filtered(px), 6)
    step = max(1, len(z) // 5000)
    png = _png_from_idats(400, 300, 3, [z[i:i + step] for i in range(0, len(z), step)])
    d = V.decode_pages([png])[0]
    assert not isinstance(d, Exception) and np.array_equal(d, px)
