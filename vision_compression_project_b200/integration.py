"""The call-site change of INTEGRATION.md §1, as code: turns the reference's per-page body into one that runs the page-image
prep on the B200.

The reference's hot path is two statements of `_process_single_page` (backend/app/pipeline/pdf_extract.py:129-130, :159):

    page_image = images[0]
    page_image.save(page_image_path)                                        # Pillow: filter + zlib + container, on the CPU
    ...
    response_text = _call_gemini_with_retry(model, EXTRACTION_PROMPT, page_image, page_num)   # the SDK encodes the image again

`patch_pdf_extract(source)` returns the module source with exactly those statements replaced (and one import added).  Nothing
else of the reference changes: thread pool, retries, JSON handling and error strings stay as they are.  The function refuses
(ValueError) when an anchor statement is not found, so a reference that moved on cannot be patched silently wrong.
tests/test_gpu_callsite.py runs the patched module with stub `pdf2image` / `google.generativeai` modules and the recorded pages.
"""
from __future__ import annotations

IMPORT_ANCHOR = "from PIL import Image\n"
IMPORT_LINE = "from vision_compression_project_b200 import prepare_page\n"

SAVE_OLD = "            page_image.save(page_image_path)\n"
SAVE_NEW = ("            prepared = prepare_page(page_image)          # convert('RGB') -> PNG -> base64 on the B200; raises like Pillow would\n"
            "            page_image_path.write_bytes(prepared.png)\n")

CALL_OLD = "_call_gemini_with_retry(model, EXTRACTION_PROMPT, page_image, page_num)"
CALL_NEW = "_call_gemini_with_retry(model, EXTRACTION_PROMPT, prepared.inline_data(), page_num)"


def patch_pdf_extract(source: str) -> str:
    """backend/app/pipeline/pdf_extract.py source -> the same source with the page-image statements on the GPU path."""
    for what, anchor in (("import block", IMPORT_ANCHOR), ("page_image.save(...)", SAVE_OLD), ("generate_content call", CALL_OLD)):
        if source.count(anchor) != 1:
            raise ValueError(f"pdf_extract.py: expected exactly one {what} anchor, found {source.count(anchor)}")
    out = source.replace(IMPORT_ANCHOR, IMPORT_ANCHOR + IMPORT_LINE, 1)
    out = out.replace(SAVE_OLD, SAVE_NEW, 1)
    return out.replace(CALL_OLD, CALL_NEW, 1)
