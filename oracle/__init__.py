"""CPU oracle for the page-image prep path (TEST INFRASTRUCTURE ONLY).

Nothing under ``oracle/`` is part of the product.  Only ``tests/``,
``__graft_entry__.smoke()`` and ``bench.py``'s CPU-baseline / ``--impl
reference`` legs may import it, and only as the checker (or as the timed CPU
baseline), never as a fallback for the CUDA path.

Two layers:

* ``oracle.pillow_path``  — the reference's own arithmetic: the Pillow calls the
  reference executes (``Image.save`` at backend/app/pipeline/pdf_extract.py:130,
  scripts/extract_pdf_with_gemini.py:152, scripts/extract_page_with_gemini.py:123)
  composed with the north-star stages (``convert`` → ``thumbnail`` size rule →
  ``resize`` → PNG → base64).  Pillow 12.2.0 ships in the image on both the
  build container and the GPU box, so this layer *is* the reference run here.
* ``oracle.restate``      — a NumPy restatement of each stage (Convert.c,
  Resample.c, Reduce.c, ZipEncode.c filter selection, PNG container, Adler-32,
  CRC-32, base64) used to localise a mismatch stage by stage.  It is pinned
  against Pillow itself and against the 23 recorded PNGs in the reference's
  ``output/`` directory (digests in tests/golden/fixtures.json).

Parity status: the reference has no tests; parity is pinned by (a) the recorded
``output/*.png`` artefacts (pixels, filter decisions, container layout) and
(b) Pillow 12.2.0 executed in the same process.  See DESIGN.md §3.
"""
