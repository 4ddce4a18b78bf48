"""Dev tool: device-only decode timing (to_device=True), several batch sizes."""
import sys, time
import numpy as np
sys.path.insert(0, ".")
import torch
import vision_compression_project_b200 as V
from vision_compression_project_b200 import synth
from tests import util as U
kind = sys.argv[1] if len(sys.argv) > 1 else "text"
pages = [synth.make_page(i, "letter", 200, photo=(kind == "photo") or (kind == "mix" and i % 4 == 3)) for i in range(8)]
ours = [r.png for r in V.prepare_pages(pages, want_base64=False)]
pil = [U.pillow_png(p) for p in pages]
for name, src in (("ours", ours), ("pillow", pil)):
    for n in (8, 64, 256):
        batch = [src[i % 8] for i in range(n)]
        V.decode_pages(batch[:8], to_device=True)
        best = 1e9
        for _ in range(3):
            torch.cuda.synchronize(); t = time.perf_counter(); d = V.decode_pages(batch, to_device=True); torch.cuda.synchronize(); best = min(best, time.perf_counter() - t)
        ok = np.array_equal(d[3].cpu().numpy(), np.asarray(pages[3]))
        print(f"decode {name} {kind} PNGs, batch {n}: {n/best:.0f} pages/s ({best*1e3:.1f} ms) ok={ok}", flush=True)
        del d
