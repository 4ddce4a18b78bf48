"""Dev tool: 8 caller threads hammer the public API with mixed page sizes / options; every result must equal the serial one."""
import sys, threading, time
import numpy as np
sys.path.insert(0, ".")
import vision_compression_project_b200 as V
from vision_compression_project_b200 import synth
jobs = []
for i in range(24):
    im = synth.make_page(i, size=(600 + 97 * (i % 5), 700 + 61 * (i % 7)), photo=(i % 3 == 0), mode="L" if i % 4 == 1 else "RGB")
    kw = [{}, {"max_side": 512}, {"mode": "L"}, {"max_side": 300, "reducing_gap": 2.0}, {"compress_level": 2}, {"compress_level": 9}][i % 6]
    jobs.append((im, kw))
ref = [V.prepare_page(im, **kw).png for im, kw in jobs]
bad = []
def work(t):
    for rep in range(6):
        for j in range(t, len(jobs), 8):
            im, kw = jobs[j]
            if V.prepare_page(im, **kw).png != ref[j]:
                bad.append((t, rep, j))
        pages = [jobs[j][0] for j in range(len(jobs)) if jobs[j][1] == {}]
        out = V.prepare_pages(pages)
        exp = [ref[j] for j in range(len(jobs)) if jobs[j][1] == {}]
        if [o.png for o in out] != exp:
            bad.append((t, rep, "batch"))
ths = [threading.Thread(target=work, args=(t,)) for t in range(8)]
t0 = time.perf_counter(); [x.start() for x in ths]; [x.join() for x in ths]
print("stress done in %.1f s, mismatches: %s" % (time.perf_counter() - t0, bad[:5] if bad else "none"))
assert not bad
