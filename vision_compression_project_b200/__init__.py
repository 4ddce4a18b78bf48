"""vcprep — B200-native page-image preparation (convert -> resize -> PNG -> base64).

The compute lives in libvcprep.so (csrc/, hand-written sm_100a CUDA behind the C ABI of include/vcprep.h);
this package is the Python host mirror of the reference's call sites.  Importing it never touches a GPU;
calling `prepare_page(s)` without the built library or without a B200 raises.
"""
from .api import decode_pages, prepare_stream  # noqa: F401
from .api import (BICUBIC, BILINEAR, BOX, HAMMING, LANCZOS, PagePrep, PreparedPage, parse_pnm, prepare_page,
                  prepare_pages, split_pnm_stream, thumbnail_size)

from .sharding import prepare_pages_all_gpus  # noqa: E402

__all__ = ["prepare_pages_all_gpus", "decode_pages", "prepare_stream", "prepare_page", "prepare_pages", "PagePrep", "PreparedPage", "thumbnail_size", "parse_pnm", "split_pnm_stream",
           "LANCZOS", "BILINEAR", "BICUBIC", "BOX", "HAMMING"]
