"""CPU: the sequential model of the GPU deflate produces valid zlib streams within the size tolerance.

The model (tests/model/deflate_model.c) is what the CUDA kernels are compared with bit-for-bit on the GPU box;
here it is checked on its own: zlib inflates its output to the input, and on the reference's recorded page it is
no larger than 1.05 x Pillow's default (level 6, Z_FILTERED) stream — the north-star's size tolerance."""
import zlib

import numpy as np
import pytest

from oracle import restate as R
from tests import model_util as M
from tests import util as U


@pytest.fixture(scope="module")
def model():
    return M.load()


def test_model_roundtrip_small_and_edge(model):
    rng = np.random.default_rng(5)
    cases = [bytes([7]), bytes(2), bytes(3), b"abc" * 1000, bytes(70000), rng.integers(0, 256, 40000, dtype=np.uint8).tobytes(),
             (rng.integers(0, 4, 600000, dtype=np.uint8) * 60).tobytes(), bytes(32768), bytes(32769), bytes(524288), bytes(524289)]
    for d in cases:
        z, st = M.deflate(model, d)
        assert zlib.decompress(z) == d, len(d)
    z, st = M.deflate(model, rng.integers(0, 256, 600000, dtype=np.uint8).tobytes())
    assert st.stored_blocks == st.blocks == 2 and len(z) <= 600000 + 5 * 11 + 6


def test_model_size_on_reference_page(model, ref_page):
    filt = R.png_filter(np.asarray(ref_page)).tobytes()
    ref = U.pillow_png(ref_page)
    z, st = M.deflate(model, filt)
    assert zlib.decompress(z) == filt
    ref_z = sum(len(c) for c in R.png_split(ref)[4])
    assert len(z) <= 1.05 * ref_z, (len(z), ref_z)
    print(f"model {len(z)} B vs Pillow {ref_z} B = {len(z) / ref_z:.3f}")
