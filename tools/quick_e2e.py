"""Dev tool: end-to-end C2 steps (pinned arrays in -> bytes out) under a few settings: stream depth, base64 on/off, group size."""
import os, sys, time
import numpy as np, torch
sys.path.insert(0, ".")
import vision_compression_project_b200 as V
from vision_compression_project_b200 import synth
if __name__ == "__main__":
    with synth.PageFactory(12) as fac:
        host = torch.empty((64, 2200, 1700, 3), dtype=torch.uint8, pin_memory=True)
        fac.arrays([(i, "letter", 200, "RGB", False) for i in range(64)], out=[host[i].numpy() for i in range(64)])
    hn = [host[i].numpy() for i in range(64)]
    def stream(depth, K=12, **kw):
        for _ in V.prepare_stream((hn for _ in range(4)), depth=depth, **kw): pass
        t = time.perf_counter(); n = 0
        for o in V.prepare_stream((hn for _ in range(K)), depth=depth, **kw): n += len(o)
        return n / (time.perf_counter() - t)
    def sync(K=8, **kw):
        for _ in range(2): V.prepare_pages(hn, **kw)
        t = time.perf_counter()
        for _ in range(K): V.prepare_pages(hn, **kw)
        return 64 * K / (time.perf_counter() - t)
    print(f"pipe_bytes={os.environ.get('VCP_PIPE_BYTES','default')} copy_threads={os.environ.get('VCP_COPY_THREADS','default')}: "
          f"sync {sync():.0f}  sync(no b64) {sync(want_base64=False):.0f}  stream d2 {stream(2):.0f} d3 {stream(3):.0f} d4 {stream(4):.0f}  "
          f"stream d3 (no b64) {stream(3, want_base64=False):.0f} pages/s", flush=True)
