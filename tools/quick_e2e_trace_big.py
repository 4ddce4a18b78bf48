"""Dev tool: VCP_TRACE timeline of one 256-page end-to-end call (steady state of the host -> device pipeline)."""
import os, sys, time
sys.path.insert(0, ".")
import numpy as np, torch
import vision_compression_project_b200 as V
from vision_compression_project_b200 import synth
if __name__ == "__main__":
    with synth.PageFactory(12) as fac:
        host = torch.empty((64, 2200, 1700, 3), dtype=torch.uint8, pin_memory=True)
        fac.arrays([(i, "letter", 200, "RGB", False) for i in range(64)], out=[host[i].numpy() for i in range(64)])
    hn = [host[i].numpy() for i in range(64)] * 4
    eng = V.PagePrep(0)
    for _ in range(2): eng.prepare_pages(hn)
    os.environ["VCP_TRACE"] = "1"
    t = time.perf_counter(); eng.prepare_pages(hn); print(f"total {1e3 * (time.perf_counter() - t):.1f} ms", file=sys.stderr)
