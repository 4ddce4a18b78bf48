"""Dev tool: single-page latency by input kind (PIL / pageable numpy / pinned numpy / device tensor) with the library's own trace."""
import sys, time, os
sys.path.insert(0, ".")
import numpy as np, torch
from PIL import Image
import vision_compression_project_b200 as V
ref = Image.open("tests/golden/ref_page_1.png"); ref.load()
a = np.ascontiguousarray(np.asarray(ref)).copy()
pin = torch.empty(a.shape, dtype=torch.uint8, pin_memory=True); pin.copy_(torch.from_numpy(a)); pn = pin.numpy()
dev = pin.cuda()
for name, src in (("PIL", ref), ("pageable numpy", a), ("pinned numpy", pn), ("device tensor", dev)):
    for _ in range(5): V.prepare_page(src)
    t = time.perf_counter()
    for _ in range(40): r = V.prepare_page(src)
    print(f"{name:22s} {(time.perf_counter()-t)/40*1e3:.2f} ms", flush=True)
os.environ["VCP_TRACE"] = "1"
for name, src in (("PIL", ref), ("pageable numpy", a)):
    print("trace", name, flush=True); V.prepare_page(src); sys.stderr.flush()
