"""Dev tool: the numbers an LZ tuning variant moves — C2 device step, its LZ stage, size, C1 device time and latency, one
synchronous e2e step, and the 5-thread single-page pattern.  Select the build with VCP_LIBRARY=<path to .so>."""
import os, sys, threading, time
import numpy as np, torch
sys.path.insert(0, ".")
import vision_compression_project_b200 as V
from vision_compression_project_b200 import _native as N, synth
from vision_compression_project_b200.api import PagePrep
from PIL import Image
if __name__ == "__main__":
    with synth.PageFactory(12) as fac:
        arrs = fac.arrays([(i, "letter", 200, "RGB", False) for i in range(64)])
    host = torch.empty((64, 2200, 1700, 3), dtype=torch.uint8, pin_memory=True)
    for i, a in enumerate(arrs): host[i] = torch.from_numpy(a)
    dev = host.cuda(); eng = PagePrep(0)
    n = 64
    descs = (N.PageDesc * n)()
    for i in range(n): descs[i].src, descs[i].width, descs[i].height, descs[i].channels = dev[i].data_ptr(), 1700, 2200, 3
    o = N.Opts(); o.out_channels, o.resample, o.compress_level, o.want_b64, o.src_device, o.dst_device = 3, 1, 6, 1, 1, 1
    bp, bb = eng.output_bound(descs, n, o)
    op = torch.empty(bp // 3, dtype=torch.uint8, device="cuda"); ob = torch.empty(bb // 3, dtype=torch.uint8, device="cuda")
    for _ in range(3): res = eng.run(descs, n, o, op.data_ptr(), bp // 3, ob.data_ptr(), bb // 3)
    torch.cuda.synchronize(); t = time.perf_counter(); lz = 0
    for _ in range(10):
        res = eng.run(descs, n, o, op.data_ptr(), bp // 3, ob.data_ptr(), bb // 3); lz += eng.stats()["ms_lz"]
    torch.cuda.synchronize(); dt = (time.perf_counter() - t) / 10
    png = sum(r.png_len for r in res) / n
    hn = [host[i].numpy() for i in range(n)]
    for _ in range(2): eng.prepare_pages(hn)
    t = time.perf_counter()
    for _ in range(5): eng.prepare_pages(hn)
    e2e = 5 * n / (time.perf_counter() - t)
    ref = Image.open("tests/golden/ref_page_1.png"); ref.load()
    for _ in range(3): r1 = V.prepare_page(ref)
    t = time.perf_counter()
    for _ in range(20): r1 = V.prepare_page(ref)
    c1 = (time.perf_counter() - t) / 20 * 1e3
    pages = [Image.fromarray(a, "RGB") for a in arrs]
    def five(reps):
        def work(k):
            for _ in range(reps):
                for i in range(k, n, 5): V.prepare_page(pages[i])
        th = [threading.Thread(target=work, args=(k,)) for k in range(5)]
        t0 = time.perf_counter(); [x.start() for x in th]; [x.join() for x in th]
        return reps * n / (time.perf_counter() - t0)
    five(1); f5 = five(3)
    print(f"{os.environ.get('VCP_LIBRARY', 'default'):40s} step {dt*1e3:6.2f} ms ({n/dt:6.0f} p/s)  lz {lz/10:5.2f} ms  png/page {png:8.0f}  "
          f"e2e sync {e2e:5.0f} p/s  C1 latency {c1:5.2f} ms ({len(r1.png)} B)  5-thread {f5:5.0f} p/s", flush=True)
