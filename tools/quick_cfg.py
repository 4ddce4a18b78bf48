"""Dev tool: per-stage device times for the BASELINE configs other than C2 (device-resident in, device out).
usage: quick_cfg.py c3|c5 [npages]"""
import sys
import time

import numpy as np
import torch

sys.path.insert(0, ".")
from vision_compression_project_b200 import _native as N, synth
from vision_compression_project_b200.api import PagePrep, _as_source

cfg = sys.argv[1]
n = int(sys.argv[2]) if len(sys.argv) > 2 else 32
if cfg == "c3":
    ims = [synth.make_page(i, "letter", 300, photo=(i % 4 == 3)) for i in range(min(n, 8))]
    kw = dict(max_side=1568, reducing_gap=None)
else:
    types = synth.mixed_page_types()
    ims = [synth.make_page(i, p, d, m, c) for i, (p, d, m, c) in enumerate(types[:min(n, 24)])]
    kw = dict(max_side=1568, reducing_gap=2.0)
uniq = [torch.from_numpy(np.asarray(im).copy()).cuda() for im in ims]
dev = [uniq[i % len(uniq)] for i in range(n)]
e = PagePrep(0)
descs = (N.PageDesc * n)()
for i, t in enumerate(dev):
    d = PagePrep._plan(_as_source(t, None), None, kw["max_side"], "RGB", 1, kw["reducing_gap"])
    descs[i] = d
o = N.Opts(); o.out_channels = 3; o.resample = 1; o.compress_level = 6; o.want_b64 = 1; o.src_device = 1; o.dst_device = 1
bp, bb = e.output_bound(descs, n, o)
op = torch.empty(bp, dtype=torch.uint8, device="cuda"); ob = torch.empty(bb, dtype=torch.uint8, device="cuda")
for it in range(4):
    torch.cuda.synchronize(); t0 = time.perf_counter()
    res = e.run(descs, n, o, op.data_ptr(), bp, ob.data_ptr(), bb)
    torch.cuda.synchronize(); dt = time.perf_counter() - t0
    st = e.stats()
    print(f"{cfg} iter {it}: {dt*1e3:.2f} ms wall, {n/dt:.0f} pages/s | " + " ".join(f"{k}={v:.2f}" for k, v in st.items() if k.startswith("ms_")) +
          f" | launches={st['kernel_launches']} in={st['in_bytes']/n/1e6:.1f} MB/page png={st['png_bytes']/n:.0f} B/page")
