"""Dev tool: is the end-to-end rate of streamed 64-page batches limited by the hand-over between batches?  One call with many pages
(the same pinned arrays repeated) never drains its pipeline between batches: its rate is what a cross-batch pipeline could reach."""
import sys, time
import numpy as np, torch
sys.path.insert(0, ".")
import vision_compression_project_b200 as V
from vision_compression_project_b200 import synth
if __name__ == "__main__":
    with synth.PageFactory(12) as fac:
        host = torch.empty((64, 2200, 1700, 3), dtype=torch.uint8, pin_memory=True)
        fac.arrays([(i, "letter", 200, "RGB", False) for i in range(64)], out=[host[i].numpy() for i in range(64)])
    hn = [host[i].numpy() for i in range(64)]
    # raw link rate
    dev = torch.empty_like(host, device="cuda")
    for _ in range(2): dev.copy_(host, non_blocking=True)
    torch.cuda.synchronize(); t = time.perf_counter()
    for _ in range(5): dev.copy_(host, non_blocking=True)
    torch.cuda.synchronize(); dt = (time.perf_counter() - t) / 5
    print(f"copy only: {dt * 1e3:.2f} ms per 64 pages = {host.numel() / dt / 1e9:.1f} GB/s -> floor {64 / dt:.0f} pages/s", flush=True)
    del dev
    for reps in (1, 4, 8):
        pages = hn * reps
        for _ in range(2): V.prepare_pages(pages)
        t = time.perf_counter(); K = max(2, 8 // reps)
        for _ in range(K): r = V.prepare_pages(pages)
        dt = (time.perf_counter() - t) / K
        print(f"one call of {len(pages)} pages: {len(pages) / dt:.0f} pages/s ({dt * 1e3:.1f} ms)", flush=True)
    for depth in (2, 3):
        for _ in V.prepare_stream((hn for _ in range(4)), depth=depth): pass
        t = time.perf_counter(); n = 0
        for o in V.prepare_stream((hn for _ in range(16)), depth=depth): n += len(o)
        print(f"stream of 64-page batches, depth {depth}: {n / (time.perf_counter() - t):.0f} pages/s", flush=True)
