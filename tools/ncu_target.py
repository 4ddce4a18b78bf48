"""Profiling target: one device-resident launch set per page class, so that an ncu capture sees every kernel of the path once with
realistic sizes.  usage: ncu_target.py [c2|c3|c5|dec|dec8p|dec64t|all]   (C2 = 64 letter-200 text pages; C3 = 16 letter-300 pages -> 1568;
C5 = 8 pages incl. 600-DPI (reduce) and an RGBA page (convert))."""
import sys
import numpy as np, torch
sys.path.insert(0, ".")
from vision_compression_project_b200 import _native as N, synth
from vision_compression_project_b200.api import PagePrep, _as_source

def run(eng, tensors, kw, reps=2):
    n = len(tensors)
    descs = (N.PageDesc * n)()
    for i, t in enumerate(tensors):
        descs[i] = PagePrep._plan(_as_source(t, None), None, kw.get("max_side"), "RGB", 1, kw.get("reducing_gap"))
    o = N.Opts(); o.out_channels, o.resample, o.compress_level, o.want_b64, o.src_device, o.dst_device = 3, 1, 6, 1, 1, 1
    bp, bb = eng.output_bound(descs, n, o)
    op = torch.empty(bp, dtype=torch.uint8, device="cuda"); ob = torch.empty(bb, dtype=torch.uint8, device="cuda")
    for _ in range(reps):
        eng.run(descs, n, o, op.data_ptr(), bp, ob.data_ptr(), bb)
    torch.cuda.synchronize()
    return eng.stats()

if __name__ == "__main__":
    what = sys.argv[1] if len(sys.argv) > 1 else "all"
    eng = PagePrep(0)
    with synth.PageFactory(12) as fac:
        if what in ("c2", "all"):
            a = fac.arrays([(i, "letter", 200, "RGB", False) for i in range(64)])
            print("c2", {k: round(v, 3) for k, v in run(eng, [torch.from_numpy(x).cuda() for x in a], {}).items() if k.startswith("ms_")})
        if what in ("c3", "all"):
            a = fac.arrays([(i, "letter", 300, "RGB", i % 4 == 3) for i in range(16)])
            print("c3", {k: round(v, 3) for k, v in run(eng, [torch.from_numpy(x).cuda() for x in a], {"max_side": 1568}).items() if k.startswith("ms_")})
        if what in ("c5", "all"):
            a = fac.arrays([(0, "letter", 600, "RGB", True), (1, "a4", 600, "L", False), (2, "legal", 600, "RGB", False), (3, "letter", 150, "L", True)])
            ts = [torch.from_numpy(x).cuda() for x in a]
            rgba = torch.from_numpy(np.concatenate([a[0], np.full(a[0].shape[:2] + (1,), 255, np.uint8)], axis=2)).cuda()
            print("c5", {k: round(v, 3) for k, v in run(eng, ts + [rgba], {"max_side": 1568, "reducing_gap": 2.0}).items() if k.startswith("ms_")})
        if what in ("dec8p", "dec64t", "dec8pp", "dec8pt"):   # decode only: 8 photo-heavy / 64 text pages, this library's PNGs; ..pp / ..pt: Pillow-written photo / text
            import vision_compression_project_b200 as V
            n, photo = (64, False) if what == "dec64t" else (8, what != "dec8pt")
            a = fac.arrays([(i, "letter", 200, "RGB", photo) for i in range(n)])
            if what in ("dec8pp", "dec8pt"):
                import io
                from PIL import Image
                ours = []
                for x in a:
                    bio = io.BytesIO(); Image.fromarray(x, "RGB").save(bio, format="PNG"); ours.append(bio.getvalue())
            else:
                ours = [r.png for r in V.prepare_pages(a, want_base64=False)]
            import time
            for _ in range(3):
                torch.cuda.synchronize(); t = time.perf_counter()
                out = eng.decode_pages(ours, to_device=True)
                torch.cuda.synchronize(); dt = time.perf_counter() - t
            assert bytes(out[3].cpu().numpy().tobytes()) == a[3].tobytes()
            print(what, "ok", f"{n / dt:.0f} pages/s ({dt * 1e3:.2f} ms)")
        if what in ("dec",):
            import io
            from PIL import Image
            import vision_compression_project_b200 as V
            a = fac.arrays([(i, "letter", 200, "RGB", i % 4 == 3) for i in range(32)])
            ours = [r.png for r in V.prepare_pages(a, want_base64=False)]
            pil = []
            for x in a[:8]:
                b = io.BytesIO(); Image.fromarray(x, "RGB").save(b, format="PNG"); pil.append(b.getvalue())
            for _ in range(2):
                out = eng.decode_pages(ours + pil, to_device=True)
            torch.cuda.synchronize()
            assert all(not isinstance(o, Exception) for o in out) and bytes(out[3].cpu().numpy().tobytes()) == a[3].tobytes()
            print("dec ok", len(out))
