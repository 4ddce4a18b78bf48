"""Dev tool: the reference's calling pattern (N threads, one page per prepare_page call) — throughput against thread count, input kind
and the combiner, to see what limits it."""
import os, sys, threading, time
import numpy as np, torch
sys.path.insert(0, ".")
import vision_compression_project_b200 as V
from vision_compression_project_b200 import synth
from PIL import Image
if __name__ == "__main__":
    with synth.PageFactory(12) as fac:
        arrs = fac.arrays([(i, "letter", 200, "RGB", False) for i in range(40)])
    pil = [Image.fromarray(a, "RGB") for a in arrs]
    host = torch.empty((40, 2200, 1700, 3), dtype=torch.uint8, pin_memory=True)
    for i, a in enumerate(arrs): host[i] = torch.from_numpy(a)
    pinned = [host[i].numpy() for i in range(40)]
    dev = [host[i].cuda() for i in range(40)]
    def run(pages, T, reps=2, **kw):
        def work(k):
            for _ in range(reps):
                for i in range(k, len(pages), T): V.prepare_page(pages[i], **kw)
        th = [threading.Thread(target=work, args=(k,)) for k in range(T)]
        t0 = time.perf_counter(); [x.start() for x in th]; [x.join() for x in th]
        return reps * len(pages) / (time.perf_counter() - t0)
    for name, pages in (("PIL", pil), ("pinned numpy", pinned), ("device tensor", dev)):
        for nb in (False, True):
            if nb: os.environ["VCP_NO_COMBINE"] = "1"
            else: os.environ.pop("VCP_NO_COMBINE", None)
            run(pages, 5, 1)
            out = []
            for T in (1, 2, 3, 5, 8):
                out.append(f"T={T}: {run(pages, T):5.0f}")
            print(f"{name:14s} {'no combiner' if nb else 'combiner   '}  " + "  ".join(out) + "  pages/s", flush=True)
    os.environ.pop("VCP_NO_COMBINE", None)
    r = run(pil, 5, 2, want_base64=False); print(f"PIL, 5 threads, no base64: {r:.0f} pages/s")
