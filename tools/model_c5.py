"""Dev tool: PNG size of the deflate model (== the kernels, bit for bit) against Pillow for the 48 page types of BASELINE config C5
(and the C3 / C4 page classes), on the CPU.  usage: model_c5.py [json of model parameter overrides]"""
import json, sys, time
from concurrent.futures import ThreadPoolExecutor
sys.path.insert(0, ".")
from oracle import restate as R
from oracle.pillow_path import prepare_page_cpu
from tests import model_util as M, util as U
from vision_compression_project_b200 import synth
lib = M.load()
kw = json.loads(sys.argv[1]) if len(sys.argv) > 1 else {}
cases = [(f"C5 {p}-{d}-{m}-{'photo' if c else 'text'}", (i, p, d, m, c), dict(max_side=1568, reducing_gap=2.0)) for i, (p, d, m, c) in enumerate(synth.mixed_page_types())]
cases += [("C3 text", (0, "letter", 300, "RGB", False), dict(max_side=1568)), ("C3 photo", (3, "letter", 300, "RGB", True), dict(max_side=1568)),
          ("C4 text", (0, "letter", 200, "RGB", False), {}), ("C4 photo", (3, "letter", 200, "RGB", True), {})]
def one(case):
    name, spec, pkw = case
    im = synth.make_page(*spec)
    ref, _, out = prepare_page_cpu(im, **pkw)
    filt = U.png_filtered(ref)
    bpp = len(out.getbands())
    z, st = M.deflate(lib, filt, bpp=bpp, rowlen=1 + out.width * bpp, **kw)
    nblk = (len(filt) + 512 * 1024 - 1) // (512 * 1024)
    ours = 8 + 25 + 12 + 12 * nblk + len(z)
    return name, ours, len(ref), out.size
worst = []
with ThreadPoolExecutor(8) as ex:
    for name, ours, ref, size in ex.map(one, cases):
        r = ours / ref
        worst.append((r, name))
        print(f"{name:34s} {size!s:14s} ours {ours:9d}  Pillow {ref:9d}  {r:.4f}{'   <-- over 1.05' if r > 1.05 else ''}", flush=True)
worst.sort(reverse=True)
print("worst:", worst[:5])
