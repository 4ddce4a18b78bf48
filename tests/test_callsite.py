"""The reference's own call site running on the replacement.

`_process_single_page` / `extract_pdf_to_page_jsons` (backend/app/pipeline/pdf_extract.py:71-207, :210-364) are loaded UNMODIFIED from
baseline/_ref/backend (copied there by tests/golden/fetch_recorded.py; /root/reference does not exist on the GPU box), with the two
network/rasteriser dependencies stubbed as the north-star says ("stubbed with recorded responses"):
  * `pdf2image.convert_from_path`  -> the recorded Poppler pages (tests/golden/_recorded/pages/page_0NN.png, decoded)
  * `google.generativeai`          -> a model whose `generate_content` returns the recorded response of that page
                                      (output/pages/page_0NN.json "raw_response") and records what it was given.
CPU test: the harness itself — the unpatched reference writes Pillow's PNG.  GPU test: the INTEGRATION.md patch
(vision_compression_project_b200.integration.patch_pdf_extract) is applied to the source text and the same calls must write PNGs that decode to the
recorded pixels, hand the SDK `PreparedPage.inline_data()`, and leave every other behaviour (JSON files, stats, error strings) as it was.
"""
import hashlib
import io
import json
import os
import sys
import threading
import types

import pytest
from PIL import Image

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF_BACKEND = os.path.join(ROOT, "baseline", "_ref", "backend")
SITE = os.path.join(REF_BACKEND, "app", "pipeline", "pdf_extract.py")
REC = os.path.join(ROOT, "tests", "golden", "_recorded", "pages")

needs_site = pytest.mark.skipif(not (os.path.exists(SITE) and os.path.exists(os.path.join(REC, "page_001.json"))),
                                reason="reference call site not fetched (python tests/golden/fetch_recorded.py in the build container)")


def _sha(b):
    return hashlib.sha256(b).hexdigest()[:16]


class _Stubs:
    """sys.modules entries for pdf2image and google.generativeai, plus the log of what the model was called with."""

    def __init__(self, n_pages):
        self.calls = []            # (page_num, parts) per generate_content call
        self.lock = threading.Lock()
        self.n_pages = n_pages
        stubs = self

        def convert_from_path(pdf_path, first_page=None, last_page=None, dpi=200, **kw):
            first = first_page or 1
            last = min(last_page or stubs.n_pages, stubs.n_pages)
            out = []
            for p in range(first, last + 1):
                im = Image.open(os.path.join(REC, f"page_{p:03d}.png"))
                im.load()
                im.info["page"] = p
                out.append(im)
            return out

        class _Response:
            def __init__(self, text):
                self.text = text

        class GenerativeModel:
            def __init__(self, model_name=None, generation_config=None):
                self.model_name = model_name

            def generate_content(self, parts):
                image = parts[1]
                # which page is this?  the stub keys the recorded response on the image handed over
                if isinstance(image, dict):
                    im = Image.open(io.BytesIO(image["data"])); im.load()
                else:
                    im = image
                page = stubs.page_of[_sha(im.tobytes())]
                with stubs.lock:
                    stubs.calls.append((page, parts))
                rec = json.load(open(os.path.join(REC, f"page_{page:03d}.json"), encoding="utf-8"))
                return _Response(rec.get("raw_response") or json.dumps(rec))

        genai = types.ModuleType("google.generativeai")
        genai.configure = lambda **kw: None
        genai.GenerativeModel = GenerativeModel
        google = types.ModuleType("google")
        google.generativeai = genai
        pdf2image = types.ModuleType("pdf2image")
        pdf2image.convert_from_path = convert_from_path
        self.modules = {"google": google, "google.generativeai": genai, "pdf2image": pdf2image}
        self.page_of = {}
        for p in range(1, n_pages + 1):
            im = Image.open(os.path.join(REC, f"page_{p:03d}.png")); im.load()
            self.page_of[_sha(im.tobytes())] = p


def _load_call_site(monkeypatch, stubs, patched):
    """Import the reference's pdf_extract.py (optionally through the INTEGRATION.md patch) as a fresh module."""
    monkeypatch.setenv("GEMINI_API_KEY", "stub-key")
    for name, mod in stubs.modules.items():
        monkeypatch.setitem(sys.modules, name, mod)
    for name in [m for m in sys.modules if m == "app" or m.startswith("app.")]:
        monkeypatch.delitem(sys.modules, name)
    monkeypatch.syspath_prepend(REF_BACKEND)
    source = open(SITE, encoding="utf-8").read()
    if patched:
        from vision_compression_project_b200.integration import patch_pdf_extract
        source = patch_pdf_extract(source)
    mod = types.ModuleType("app.pipeline.pdf_extract_under_test")
    mod.__file__ = SITE
    exec(compile(source, SITE, "exec"), mod.__dict__)
    return mod


@needs_site
def test_patch_applies_to_the_reference_source():
    from vision_compression_project_b200.integration import CALL_NEW, SAVE_NEW, patch_pdf_extract
    src = open(SITE, encoding="utf-8").read()
    out = patch_pdf_extract(src)
    assert SAVE_NEW in out and CALL_NEW in out and "page_image.save(" not in out
    assert len(out.splitlines()) == len(src.splitlines()) + 2            # one import, one extra statement; nothing else moved
    with pytest.raises(ValueError):
        patch_pdf_extract(out)                                          # anchors are gone: refuses to patch twice


@needs_site
def test_unpatched_reference_with_stubs_writes_pillows_png(monkeypatch, tmp_path, fixtures):
    """The harness on the CPU: the reference as it is, with the stubs, reproduces the recorded pixels and gets the recorded response."""
    stubs = _Stubs(2)
    mod = _load_call_site(monkeypatch, stubs, patched=False)
    images, pages = tmp_path / "images", tmp_path / "pages"
    images.mkdir(); pages.mkdir()
    ok, err, data = mod._process_single_page(1, tmp_path / "doc.pdf", 200, images, pages)
    assert ok and err is None and data["page_number"] == 1 and data["markdown"]
    dec = Image.open(images / "page_001.png"); dec.load()
    assert _sha(dec.tobytes()) == fixtures["fixtures"]["pages/page_001.png"]["sha_px"]
    assert isinstance(stubs.calls[0][1][1], Image.Image)                # the SDK is handed the PIL image (and would encode it again)


@needs_site
@pytest.mark.gpu
def test_reference_call_site_on_the_gpu_path(monkeypatch, tmp_path, fixtures):
    import vision_compression_project_b200 as V
    stubs = _Stubs(6)
    mod = _load_call_site(monkeypatch, stubs, patched=True)
    images, pages = tmp_path / "images", tmp_path / "pages"
    images.mkdir(); pages.mkdir()
    # ---- one page, as _process_single_page is called by the pool
    ok, err, data = mod._process_single_page(1, tmp_path / "doc.pdf", 200, images, pages)
    assert ok and err is None, err
    png = (images / "page_001.png").read_bytes()
    fx = fixtures["fixtures"]["pages/page_001.png"]
    dec = Image.open(io.BytesIO(png)); dec.load()
    assert dec.mode == "RGB" and dec.size == tuple(fx["size"]) and _sha(dec.tobytes()) == fx["sha_px"]
    assert len(png) <= 1.05 * fx["pillow_png_bytes"] and len(png) <= 1.05 * fx["bytes"]
    page, parts = stubs.calls[0]
    rec_im = Image.open(os.path.join(REC, "page_001.png")); rec_im.load()
    assert page == 1 and parts[0] == mod.EXTRACTION_PROMPT
    assert parts[1] == V.prepare_page(rec_im).inline_data() == {"mime_type": "image/png", "data": png}
    # everything behind the image is untouched: the page JSON is byte for byte what the UNPATCHED reference writes from the same response
    stubs0 = _Stubs(1)
    mod0 = _load_call_site(monkeypatch, stubs0, patched=False)
    images0, pages0 = tmp_path / "images0", tmp_path / "pages0"
    images0.mkdir(); pages0.mkdir()
    ok0, err0, _ = mod0._process_single_page(1, tmp_path / "doc.pdf", 200, images0, pages0)
    assert ok0 and (pages / "page_001.json").read_bytes() == (pages0 / "page_001.json").read_bytes()
    assert json.load(open(pages / "page_001.json", encoding="utf-8"))["page_number"] == 1
    for name, m_ in stubs.modules.items():              # the second load swapped the stub modules: put this test's back
        monkeypatch.setitem(sys.modules, name, m_)
    # ---- the whole document through the reference's 5-thread pool (pdf_extract.py:313-350)
    stats = mod.extract_pdf_to_page_jsons(tmp_path / "doc.pdf", pages, images, dpi=200, overwrite=True)
    assert stats == {"pages_total": 6, "processed_pages": [1, 2, 3, 4, 5, 6], "failed_pages": []}
    for p in range(1, 7):
        fxp = fixtures["fixtures"][f"pages/page_{p:03d}.png"]
        d = Image.open(images / f"page_{p:03d}.png"); d.load()
        assert _sha(d.tobytes()) == fxp["sha_px"], p
        assert (pages / f"page_{p:03d}.json").exists()
    assert sorted(c[0] for c in stubs.calls[1:]) == [1, 2, 3, 4, 5, 6]
    assert all(isinstance(c[1][1], dict) and c[1][1]["mime_type"] == "image/png" for c in stubs.calls)
    # ---- the reference's error convention survives: a page the GPU path rejects becomes a failed page, not an exception
    mod.convert_from_path = lambda *a, **k: [Image.new("CMYK", (8, 8))]         # (the module bound the name at import)
    ok, err, data = mod._process_single_page(3, tmp_path / "doc.pdf", 200, images, pages, overwrite=True)
    assert not ok and err.startswith("Error converting page 3 to image: ValueError")


def test_inline_data_helper():
    from vision_compression_project_b200 import PreparedPage
    p = PreparedPage(b"\x89PNG...", b"iVBORw==", (1, 1), "RGB")
    assert p.inline_data() == {"mime_type": "image/png", "data": b"\x89PNG..."}
    assert p.inline_data(base64_text=True) == {"mime_type": "image/png", "data": "iVBORw=="}
    assert p.blob() == ("image/png", b"\x89PNG...")
    with pytest.raises(ValueError):
        PreparedPage(None, None, (0, 0), "", error="ValueError: bad").inline_data()
    with pytest.raises(ValueError):
        PreparedPage(b"x", None, (1, 1), "RGB").inline_data(base64_text=True)
