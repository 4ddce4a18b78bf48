"""Build libvcprep.so (sm_100a only) in-tree with nvcc.  `python -m vision_compression_project_b200.build`

The shared library is the product's only compute path: there is no CPU fallback, and importing the
package on a machine where the library has not been built raises immediately (see _native.py).
"""
from __future__ import annotations

import os
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
SOURCES = ["api.cu", "pixel_ops.cu", "png_filter.cu", "deflate_lz.cu", "deflate_huff.cu", "png_container.cu", "png_decode.cu"]
LIB = os.environ.get("VCP_LIBRARY") or os.path.join(HERE, "libvcprep.so")    # VCP_LIBRARY: dev A/B builds


def nvcc_path() -> str:
    for cand in (os.environ.get("NVCC"), shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError("nvcc not found (set NVCC=...)")


def needs_build() -> bool:
    if not os.path.exists(LIB):
        return True
    t = os.path.getmtime(LIB)
    deps = [os.path.join(CSRC, f) for f in os.listdir(CSRC)] + [os.path.join(HERE, "..", "include", "vcprep.h")]
    return any(os.path.getmtime(d) > t for d in deps if os.path.exists(d))


def build(force: bool = False, verbose: bool = False) -> str:
    if not force and not needs_build():
        return LIB
    cmd = [nvcc_path(), "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
           "-Xcompiler", "-fPIC", "-shared", "-o", LIB + ".tmp"] + [os.path.join(CSRC, s) for s in SOURCES]
    cmd += os.environ.get("VCP_NVCC_FLAGS", "").split()          # dev: -D overrides of tuning constants for A/B builds
    if verbose:
        cmd += ["-Xptxas", "-v"]
        print(" ".join(cmd), flush=True)
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError("nvcc failed:\n" + r.stdout + r.stderr)
    if verbose:
        print(r.stdout + r.stderr)
    os.replace(LIB + ".tmp", LIB)
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
