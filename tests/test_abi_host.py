"""CPU: the C-ABI library loads and exports exactly what include/vcprep.h declares; host-side logic
(planning helpers, Pillow's coefficient tables computed inside the .so, PNM parsing, error mapping)."""
import ctypes as C
import os
import re
import subprocess

import numpy as np
import pytest
from PIL import Image

import vision_compression_project_b200 as V
from oracle import pillow_path as PP
from oracle import restate as R
from vision_compression_project_b200 import _native as N
from vision_compression_project_b200 import build as B

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def lib():
    B.build()
    return N.load()


def header_symbols():
    src = open(os.path.join(ROOT, "include", "vcprep.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(vcp_[a-z0-9_]+)\s*\(", src)))


def test_library_exports_every_declared_symbol(lib):
    declared = header_symbols()
    assert declared == sorted(N.SYMBOLS), "ctypes table and header disagree"
    out = subprocess.run(["nm", "-D", "--defined-only", N.LIB_PATH], capture_output=True, text=True, check=True).stdout
    exported = set(re.findall(r" T (vcp_[a-z0-9_]+)", out))
    assert set(declared) <= exported, sorted(set(declared) - exported)
    header = open(os.path.join(ROOT, "include", "vcprep.h")).read()
    assert lib.vcp_version() == int(re.search(r"#define VCP_VERSION (\d+)", header).group(1)) == 101


def test_library_is_sm100a_only():
    out = subprocess.run(["cuobjdump", "-lelf", N.LIB_PATH], capture_output=True, text=True).stdout
    archs = set(re.findall(r"sm_(\d+a?)", out))
    assert archs == {"100a"}, archs


def test_no_gpu_fails_loudly(lib):
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    with pytest.raises(RuntimeError):
        V.prepare_page(Image.new("RGB", (8, 8)))


def test_output_bound_and_bad_pages(lib):
    descs = (N.PageDesc * 3)()
    buf = np.zeros(100 * 80 * 3, np.uint8)
    for d in descs:
        d.src, d.width, d.height, d.channels = buf.ctypes.data, 100, 80, 3
    descs[1].channels = 7                                  # bad page: contributes nothing, others unaffected
    o = N.Opts(); o.out_channels = 3; o.want_b64 = 1; o.compress_level = 6
    p, b = C.c_uint64(), C.c_uint64()
    assert lib.vcp_output_bound(descs, 3, C.byref(o), C.byref(p), C.byref(b)) == 0
    filt = 80 * (1 + 300)
    assert 2 * filt < p.value < 2 * (filt + 200) + 64 and b.value >= 4 * p.value // 3 - 64
    assert lib.vcp_output_bound(None, 1, C.byref(o), C.byref(p), C.byref(b)) == N.VCP_EINVAL
    assert b"bad arguments" in lib.vcp_last_error()


@pytest.mark.parametrize("flt", [1, 2, 3, 4, 5])
def test_coefficients_match_restatement(lib, flt):
    cases = [(2550, 1212, None), (3300, 1568, None), (1654, 1109, None), (96, 131, None), (7, 1, None), (1, 9, None),
             (1275, 1212, (0.0, 1275.0)), (851, 400, (0.0, 850.3333))]
    for in_size, out_size, box in cases:
        b0, b1 = (0.0, float(in_size)) if box is None else box
        ks = C.c_int()
        assert lib.vcp_resample_coeffs(in_size, out_size, flt, b0, b1, None, None, C.byref(ks)) == 0
        bounds = np.zeros((out_size, 2), np.int32); kk = np.zeros((out_size, ks.value), np.int32)
        assert lib.vcp_resample_coeffs(in_size, out_size, flt, b0, b1, bounds.ctypes.data, kk.ctypes.data, C.byref(ks)) == 0
        k2, b2, kk2 = R.resample_coeffs(in_size, out_size, flt, None if box is None else box)
        assert ks.value == k2 and np.array_equal(bounds, b2) and np.array_equal(kk, kk2), (in_size, out_size, flt)


def test_thumbnail_size_matches_pillow():
    for (w, h) in [(2550, 3300), (1700, 2200), (1654, 2339), (2550, 4200), (5100, 6600), (4960, 7016), (100, 3000), (3000, 7), (33, 33)]:
        for m in (1568, 1024, 64, 5000):
            im = Image.new("L", (w, h)); im.thumbnail((m, m))
            assert V.thumbnail_size((w, h), (m, m)) == im.size == PP.thumbnail_size((w, h), (m, m))


def test_parse_pnm():
    body = bytes(range(24))
    assert V.parse_pnm(b"P6\n4 2\n255\n" + body) == (4, 2, 3, 11)
    assert V.parse_pnm(b"P5 8 3 255\n" + body) == (8, 3, 1, 11)
    assert V.parse_pnm(b"P6\n# made by pdftoppm\n4 2\n255\n" + body)[:3] == (4, 2, 3)
    for bad in (b"P3\n1 1\n255\n0 0 0", b"P6\n4 2\n65535\n" + body, b"P6\n4 4\n255\n" + body, b"hello"):
        with pytest.raises(ValueError):
            V.parse_pnm(bad)


def test_split_pnm_stream():
    a = b"P6\n4 2\n255\n" + bytes(range(24))
    b = b"P5\n3 3\n255\n" + bytes(range(9))
    parts = V.split_pnm_stream(a + b + b"\n" + a)
    assert [bytes(p) for p in parts] == [a, b, a]
    assert V.split_pnm_stream(b"") == []
    with pytest.raises(ValueError):
        V.split_pnm_stream(a + b"garbage")


def test_input_normalisation_and_planning():
    from vision_compression_project_b200.api import PagePrep, _as_source
    s = _as_source(Image.new("RGB", (5, 4)), None)                        # Pillow's own RGBX storage, zero copy (Arrow capsule)
    assert (s.w, s.h, s.c, s.logical_c, s.stride, s.device) == (5, 4, 4, 3, 20, False)
    s = _as_source(Image.new("L", (5, 4)), None)
    assert (s.w, s.h, s.c, s.logical_c, s.stride) == (5, 4, 1, 1, 5)
    s = _as_source(Image.new("LA", (5, 4)), None)                         # packed copy
    assert (s.w, s.h, s.c, s.stride) == (5, 4, 2, 10)
    px = np.arange(5 * 4 * 3, dtype=np.uint8).reshape(4, 5, 3)
    s = _as_source(Image.fromarray(px, "RGB"), None)
    import ctypes
    raw = np.ctypeslib.as_array((ctypes.c_uint8 * (5 * 4 * 4)).from_address(s.ptr)).reshape(4, 5, 4)
    assert np.array_equal(raw[:, :, :3], px)
    s = _as_source(np.zeros((4, 10, 3), np.uint8)[:, :5], None)           # row-strided view is passed through
    assert (s.w, s.h, s.c, s.stride) == (5, 4, 3, 30)
    s = _as_source(b"P5\n3 2\n255\n" + bytes(6), None)
    assert (s.w, s.h, s.c) == (3, 2, 1)
    s = _as_source(bytes(24), (2, 4, 3))
    assert (s.w, s.h, s.c) == (4, 2, 3)
    for bad in (Image.new("CMYK", (2, 2)), Image.new("1", (2, 2)), np.zeros((2, 2), np.float32), 3.5):
        with pytest.raises((ValueError, TypeError)):
            _as_source(bad, None)
    src = _as_source(Image.new("RGB", (5100, 6600)), None)
    d = PagePrep._plan(src, None, 1568, "RGB", 1, 2.0)
    assert (d.dst_width, d.dst_height, d.reduce_x, d.reduce_y) == (1212, 1568, 2, 2)
    d = PagePrep._plan(src, None, 1568, "RGB", 1, None)
    assert (d.dst_width, d.dst_height, d.reduce_x, d.reduce_y) == (1212, 1568, 0, 0)
    d = PagePrep._plan(src, None, 9000, "RGB", 1, None)                   # thumbnail never enlarges
    assert (d.dst_width, d.dst_height) == (0, 0)
    with pytest.raises(ValueError):
        PagePrep._plan(src, (0, 5), None, "RGB", 1, None)


def test_host_scatter_fills_fresh_bytes(lib):
    rng = np.random.default_rng(3)
    src = rng.integers(0, 256, 6_000_000, dtype=np.uint8)
    ranges = [(0, 0), (5, 1), (100, 1_500_000), (2_000_000, 2_999_999), (7, 3), (5_999_999, 1)]
    for threads in (1, 4, 64):
        out = N.gather_bytes(src.ctypes.data, ranges, threads)
        assert all(type(o) is bytes and o == src[a:a + n].tobytes() for o, (a, n) in zip(out, ranges))
    assert N.gather_bytes(src.ctypes.data, [], 4) == []


def test_pillow_row_table_reads_the_real_pixels():
    """PIL images above Pillow's 16 MB block size live in several blocks; api._pil_row_table finds libImaging's row-pointer table
    (validated, not assumed).  Every row read through it must be the image's row — checked here on the CPU for multi-block and
    single-block images of every storage kind; a layout that is not recognised must give None, never a wrong table."""
    import ctypes as C
    import numpy as np
    from PIL import Image
    from vision_compression_project_b200 import api
    rng = np.random.default_rng(4)
    for mode, (w, h) in (("RGB", (2550, 3300)), ("L", (5100, 4000)), ("RGBA", (2300, 2100)), ("RGB", (64, 48)), ("L", (33, 7))):
        bands = len(mode)
        a = rng.integers(0, 256, (h, w, bands), dtype=np.uint8)
        im = Image.fromarray(a[:, :, 0] if bands == 1 else a, mode)
        rt = api._pil_row_table(im)
        assert rt is not None, (mode, w, h)
        keep, table, px = rt
        rows = (C.c_void_p * h).from_address(table)
        for y in (0, 1, h // 3, h // 2, h - 2, h - 1):
            raw = np.frombuffer((C.c_ubyte * (w * px)).from_address(rows[y]), np.uint8).reshape(w, px)
            assert np.array_equal(raw[:, :bands], a[y].reshape(w, bands)), (mode, y)
        src = api._as_source(im, None)
        assert (src.row_ptrs is not None) == (api._pil_zero_copy(im) is None) and src.logical_c == bands
    assert api._pil_row_table(Image.new("CMYK", (8, 8))) is None
    assert api._pil_row_table(Image.new("LA", (8, 8))) is None
