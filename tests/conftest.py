import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a B200 (run with -m gpu on the GPU box)")


@pytest.fixture(scope="session")
def golden_dir():
    return os.path.join(ROOT, "tests", "golden")


@pytest.fixture(scope="session")
def crops(golden_dir):
    import numpy as np
    return dict(np.load(os.path.join(golden_dir, "crops.npz")))


@pytest.fixture(scope="session")
def fixtures(golden_dir):
    import json
    return json.load(open(os.path.join(golden_dir, "fixtures.json")))


@pytest.fixture(scope="session")
def ref_page(golden_dir):
    """BASELINE.json configs[0]: the reference's recorded output/page_1.png (committed golden copy)."""
    from PIL import Image
    im = Image.open(os.path.join(golden_dir, "ref_page_1.png"))
    im.load()
    return im
