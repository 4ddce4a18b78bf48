"""CPU: the hand-filtered PNGs the GPU un-filter test feeds the decoder are what they claim to be — Pillow decodes them to the
pixels they were made from, for every channel count and filter mix (so a GPU failure there is the kernel's, not the generator's)."""
import io

import numpy as np
import pytest
from PIL import Image

from tests.test_gpu_decode import _filtered_png


@pytest.mark.parametrize("c", [1, 2, 3, 4])
def test_hand_filtered_png_is_a_valid_png_of_its_pixels(c):
    rng = np.random.default_rng(c)
    for h, w, mix in [(1, 1, [0]), (7, 33, [1]), (40, 5, [2]), (33, 64, [3]), (65, 31, [4]), (50, 67, [0, 1, 2, 3, 4]), (37, 100, [4, 3])]:
        px = rng.integers(0, 256, (h, w, c), dtype=np.uint8)
        im = Image.open(io.BytesIO(_filtered_png(px, mix)))
        im.load()
        assert im.size == (w, h) and im.mode == {1: "L", 2: "LA", 3: "RGB", 4: "RGBA"}[c]
        assert np.array_equal(np.asarray(im).reshape(h, w, c), px), (h, w, c, mix)
