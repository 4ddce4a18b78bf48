"""Dev tool: LZ/filter stage time for homogeneous batches (blank / text / photo / noise) to see per-path costs."""
import sys, time
import numpy as np, torch
sys.path.insert(0, ".")
from vision_compression_project_b200 import _native as N, synth
from vision_compression_project_b200.api import PagePrep
n = 64
e = PagePrep(0)
kinds = {
    "blank": lambda i: np.full((2200, 1700, 3), 255, np.uint8),
    "text": lambda i: np.asarray(synth.make_page(i, "letter", 200)),
    "photo": lambda i: np.asarray(synth.make_page(i, "letter", 200, photo=True)),
    "noise": lambda i: np.random.default_rng(i).integers(0, 256, (2200, 1700, 3), dtype=np.uint8),
    "smooth": lambda i: (np.add.outer(np.arange(2200) // 3, np.arange(1700) // 2)[:, :, None] + np.array([0, 40, 90])).astype(np.uint8),
}
only = sys.argv[1].split(',') if len(sys.argv) > 1 else list(kinds)
for name, fn in kinds.items():
    if name not in only:
        continue
    uniq = [torch.from_numpy(fn(i).copy()).cuda() for i in range(4)]
    dev = [uniq[i % 4] for i in range(n)]
    descs = (N.PageDesc * n)()
    for d, t in zip(descs, dev):
        d.src, d.width, d.height, d.channels = t.data_ptr(), t.shape[1], t.shape[0], 3
    o = N.Opts(); o.out_channels = 3; o.compress_level = int(__import__('os').environ.get('LEVEL', '6')); o.want_b64 = 1; o.src_device = 1; o.dst_device = 1
    bp, bb = e.output_bound(descs, n, o)
    op = torch.empty(bp, dtype=torch.uint8, device="cuda"); ob = torch.empty(bb, dtype=torch.uint8, device="cuda")
    for it in range(3):
        e.run(descs, n, o, op.data_ptr(), bp, ob.data_ptr(), bb)
    st = e.stats()
    print(f"{name:7s} " + " ".join(f"{k}={v:.2f}" for k, v in st.items() if k.startswith("ms_") and v > 0.005) + f" png={st['png_bytes']/n:.0f} B/page", flush=True)
    del op, ob
