"""CPU: the oracle restatement (oracle/restate.py) against Pillow itself and the committed golden vectors.

The golden vectors come from the reference's recorded artefacts (output/*.png) and from Pillow 12.2.0 run on
them (tests/golden/make_golden.py).  They pin: pixels, filter decisions, Adler-32, container layout.
"""
import base64
import io
import os
import zlib

import numpy as np
import pytest
from PIL import Image

from oracle import pillow_path as PP
from oracle import restate as R
from tests import util as U

FLT = {"lanczos": 1, "bilinear": 2, "bicubic": 3, "box": 4, "hamming": 5}


def test_fixture_page1_matches_recorded_digests(ref_page, fixtures):
    fx = fixtures["fixtures"]["page_1.png"]
    assert ref_page.size == tuple(fx["size"]) and ref_page.mode == fx["mode"]
    assert U.sha(ref_page.tobytes()) == fx["sha_px"]
    raw = open(os.path.join(os.path.dirname(__file__), "golden", "ref_page_1.png"), "rb").read()
    assert U.sha(raw) == fx["sha_file"] and len(raw) == fx["bytes"]
    w, h, bd, ct, idat, ok = R.png_split(raw)
    assert ok and ct == fx["color_type"] and len(idat) == fx["n_idat"] and max(map(len, idat)) == fx["idat_max"]
    filt = zlib.decompress(b"".join(idat))
    assert U.sha(filt) == fx["sha_filtered"] and f"{zlib.adler32(filt):08x}" == fx["adler32"]


def test_filter_restatement_reproduces_recorded_page(ref_page, fixtures):
    """2339 recorded filter decisions + residual bytes of the reference's own PNG."""
    fx = fixtures["fixtures"]["page_1.png"]
    filt = R.png_filter(np.asarray(ref_page))
    assert U.sha(filt.tobytes()) == fx["sha_filtered"]
    hist = np.bincount(filt.reshape(ref_page.height, -1)[:, 0], minlength=5).tolist()
    assert hist == fx["filter_hist_NSUAP"]
    assert R.adler32(filt) == int(fx["adler32"], 16)


@pytest.mark.skipif(not os.path.isdir("/root/reference/output/pages"), reason="reference artefacts not on this machine")
def test_filter_restatement_all_recorded_pages(fixtures):
    for name, fx in fixtures["fixtures"].items():
        im = Image.open(os.path.join("/root/reference/output", name)); im.load()
        assert U.sha(im.tobytes()) == fx["sha_px"]
        assert U.sha(R.png_filter(np.asarray(im)).tobytes()) == fx["sha_filtered"], name


def test_crops_filter_convert_resize_reduce(crops):
    for ci in range(3):
        px = crops[f"c{ci}_px"]
        assert np.array_equal(R.png_filter(px), crops[f"c{ci}_filtered"])
        assert np.array_equal(R.png_filter(px, optimize=True), crops[f"c{ci}_filtered_opt"])
        g = R.convert_mode(px, "RGB", "L")
        assert np.array_equal(g[:, :, 0], crops[f"c{ci}_L"])
        assert np.array_equal(R.png_filter(g), crops[f"c{ci}_L_filtered"])
        for fname, flt in FLT.items():
            for size in [(61, 47), (131, 117)]:
                assert np.array_equal(R.resample(px, size, flt), crops[f"c{ci}_{fname}_{size[0]}x{size[1]}"]), (ci, fname, size)
        assert np.array_equal(R.reduce_box(px, 2, 2), crops[f"c{ci}_reduce2"])
        assert np.array_equal(R.reduce_box(px, 3, 2), crops[f"c{ci}_reduce3x2"])


def test_restatement_vs_pillow_random():
    rng = np.random.default_rng(7)
    for t in range(40):
        h, w = int(rng.integers(1, 40)), int(rng.integers(1, 40))
        mode = ["L", "RGB", "RGBA"][t % 3]
        c = R.BPP[mode]
        kind = t % 4
        if kind == 0:
            px = rng.integers(0, 256, (h, w, c), dtype=np.uint8)
        elif kind == 1:
            px = np.full((h, w, c), int(rng.integers(0, 256)), np.uint8)
        else:
            px = (np.add.outer(np.arange(h) * 3, np.arange(w) * 5)[:, :, None] + np.arange(c) * 17).astype(np.uint8)
        im = Image.fromarray(px[:, :, 0] if c == 1 else px, mode)
        for opt in (False, True):
            ref = U.png_filtered(U.pillow_png(im, optimize=opt))
            assert R.png_filter(px, optimize=opt).tobytes() == ref, (t, mode, opt)
        if mode != "RGBA":
            ow, oh = int(rng.integers(1, 50)), int(rng.integers(1, 50))
            flt = int(rng.integers(1, 6))
            ref = np.asarray(im.resize((ow, oh), flt)).reshape(oh, ow, c)
            assert np.array_equal(R.resample(px, (ow, oh), flt), ref), (t, mode, flt, (w, h), (ow, oh))
        if mode == "RGBA":
            assert np.array_equal(R.convert_mode(px, "RGBA", "RGB"), np.asarray(im.convert("RGB")))
        if mode == "L":
            assert np.array_equal(R.convert_mode(px, "L", "RGB"), np.asarray(im.convert("RGB")))
        if mode == "RGB":
            assert np.array_equal(R.convert_mode(px, "RGB", "L")[:, :, 0], np.asarray(im.convert("L")))


def test_thumbnail_with_reducing_gap_composition():
    rng = np.random.default_rng(3)
    for (w, h, m) in [(300, 400, 40), (257, 191, 31), (640, 100, 64)]:
        px = rng.integers(0, 256, (h, w, 3), dtype=np.uint8)
        im = Image.fromarray(px, "RGB")
        tw, th = PP.thumbnail_size((w, h), (m, m))
        ref = im.resize((tw, th), Image.Resampling.LANCZOS, reducing_gap=2.0)
        fx = int(w / tw / 2.0) or 1
        fy = int(h / th / 2.0) or 1
        cur = R.reduce_box(px, fx, fy) if (fx > 1 or fy > 1) else px
        out = R.resample(cur, (tw, th), R.LANCZOS, box=(0.0, 0.0, w / fx, h / fy))
        assert np.array_equal(out, np.asarray(ref)), (w, h, m)
        t2 = im.copy(); t2.thumbnail((m, m), Image.Resampling.LANCZOS, reducing_gap=2.0)
        assert t2.size == (tw, th) and np.array_equal(np.asarray(t2), np.asarray(ref))


def test_checksums_container_base64():
    rng = np.random.default_rng(1)
    for n in [0, 1, 2, 3, 4, 5, 57, 4095, 4096, 4097, 70000]:
        d = rng.integers(0, 256, n, dtype=np.uint8).tobytes()
        assert R.adler32(d) == zlib.adler32(d)
        assert R.b64encode(d) == base64.b64encode(d)
        if n <= 4097:
            assert R.crc32(d) == zlib.crc32(d)
    px = rng.integers(0, 256, (9, 7, 3), dtype=np.uint8)
    z = zlib.compress(R.png_filter(px).tobytes())
    png = R.png_wrap(7, 9, 3, z)
    assert np.array_equal(np.asarray(Image.open(io.BytesIO(png))), px)
    assert np.array_equal(R.png_unfilter(R.png_filter(px), 9, 7, 3), px)


def test_pillow_path_default_is_plain_save(ref_page):
    crop = ref_page.crop((100, 300, 500, 420))
    png, b64, im = PP.prepare_page_cpu(crop)
    assert png == U.pillow_png(crop) and b64 == base64.b64encode(png) and im.size == crop.size
    png2, _, im2 = PP.prepare_page_cpu(crop, max_side=64)
    t = crop.copy(); t.thumbnail((64, 64), Image.Resampling.LANCZOS, reducing_gap=None)
    assert im2.size == t.size
