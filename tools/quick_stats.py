"""Dev tool: per-stage device times of one C2-like batch (device-resident in, device out)."""
import ctypes as C
import sys
import time

import numpy as np
import torch

sys.path.insert(0, ".")
from vision_compression_project_b200 import _native as N, synth
from vision_compression_project_b200.api import PagePrep

n = int(sys.argv[1]) if len(sys.argv) > 1 else 64
photo_every = int(sys.argv[2]) if len(sys.argv) > 2 else 4
uniq = [np.asarray(synth.make_page(i, "letter", 200, photo=bool(photo_every) and (i % photo_every == photo_every - 1))) for i in range(min(n, 8))]
dev = [torch.from_numpy(uniq[i % len(uniq)].copy()).cuda() for i in range(n)]
e = PagePrep(0)
descs = (N.PageDesc * n)()
for d, t in zip(descs, dev):
    d.src, d.width, d.height, d.channels = t.data_ptr(), t.shape[1], t.shape[0], 3
o = N.Opts(); o.out_channels = 3; o.compress_level = 6; o.want_b64 = 1; o.src_device = 1; o.dst_device = 1
bp, bb = e.output_bound(descs, n, o)
op = torch.empty(bp, dtype=torch.uint8, device="cuda"); ob = torch.empty(bb, dtype=torch.uint8, device="cuda")
for it in range(4):
    torch.cuda.synchronize(); t0 = time.perf_counter()
    res = e.run(descs, n, o, op.data_ptr(), bp, ob.data_ptr(), bb)
    torch.cuda.synchronize(); dt = time.perf_counter() - t0
    st = e.stats()
    print(f"iter {it}: {dt*1e3:.2f} ms wall, {n/dt:.0f} pages/s | " + " ".join(f"{k}={v:.2f}" for k, v in st.items() if k.startswith("ms_")) +
          f" | launches={st['kernel_launches']} png={st['png_bytes']/n:.0f} B/page")
