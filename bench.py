#!/usr/bin/env python
"""bench.py — pages/sec of the page-image prep path (convert -> [resize] -> PNG filter+deflate -> base64) on B200.

Contract (task statement §④):  python bench.py --gpus N --steps K --warmup W  [--impl reference]
  * headline workload at every N: BASELINE.json configs[1] per GPU — a batch of 64 synthetic letter-size pages at 200 DPI
    (1700x2200 RGB, seeds = global page index); "step" = one pass of the whole path over that batch.
    N > 1 is weak scaling: every rank owns its own page range, no data-path collective (SURVEY.md §8 e).
  * value   = pages/s with the pages already resident in HBM and the outputs left in HBM (whole job, all ranks).
  * e2e     = the same metric through the public function `prepare_stream` / `prepare_pages` with HOST (pinned) page buffers in and
              Python bytes out: H2D, every kernel, D2H and the bytes slicing are inside the timed region; next to it the same call
              with PIL images in (the reference's own input type) and the reference's calling pattern (5 threads, one page per call).
  * roofline= the dominant kernel (LZ77 match finding) — algorithmic bytes of the step / its CUDA-event time.
  * cpu_baseline / --impl reference = the reference's own CPU path (Pillow save + base64, oracle/pillow_path.py)
    on all host cores of this box, on the same 64 pages.
  * configs = the other BASELINE.json configs, each pixel-checked against the Pillow path outside the timed region:
      C1 the reference's recorded output/page_1.png, one call;  C3 256 letter-300 pages -> LANCZOS 1568;  C5 the 48-type mix;
      S150 the service's default 150-DPI pages  (N = 1)
      c4  = the 2,000-page document (25 % photo pages) sharded by page range over the N ranks through sharding.py: STRONG scaling.
            A rank consumes its results batch by batch (the reference's workers write a page and drop it); every rank pins itself to
            its share of the host cores and keeps freed result memory in its heap (sharding.pin_rank_to_cores / keep_result_memory).
"""
from __future__ import annotations

import argparse
import base64
import hashlib
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

PAGES_PER_GPU = 64
PAPER, DPI = "letter", 200
WORKLOAD = "C2: 64 synthetic letter-size pages @200 DPI (1700x2200 RGB) per GPU, convert('RGB') + PNG + base64"
L2_NOTE = "inputs (718 MB/step) larger than L2, no flush needed"
C3_PAGES, C4_PAGES, MAX_SIDE = 256, 2000, 1568


def host_cores() -> int:
    try:
        return len(os.sched_getaffinity(0))
    except AttributeError:  # pragma: no cover
        return os.cpu_count() or 1


def bench_config() -> dict:
    """`config` of the JSON line — the same dict for this arm and for --impl reference."""
    return {"workload": WORKLOAD, "pages_per_gpu": PAGES_PER_GPU, "l2": L2_NOTE}


# ------------------------------------------------------------------------------------------------ synthetic pages
def PageFactory(workers: int):
    from vision_compression_project_b200 import synth
    return synth.PageFactory(workers)


def c2_specs(first_seed, n):
    return [(s, PAPER, DPI, "RGB", False) for s in range(first_seed, first_seed + n)]


def c3_specs(n):
    return [(s, "letter", 300, "RGB", s % 4 == 3) for s in range(n)]


def c4_specs(lo, hi):
    return [(s, "letter", 200, "RGB", s % 4 == 3) for s in range(lo, hi)]


def c5_specs():
    from vision_compression_project_b200 import synth
    return [(i, p, d, m, c) for i, (p, d, m, c) in enumerate(synth.mixed_page_types())]


def to_pil(a):
    from PIL import Image
    return Image.fromarray(a, "RGB" if a.ndim == 3 else "L")


# ------------------------------------------------------------------------------------------------ CPU arm
def cpu_path_pages_per_s(pages, threads: int, repeats: int = 1, **kw):
    """The reference's CPU path (Pillow [resize +] PNG save + base64) over `pages` on `threads` host threads.
    Pillow releases the GIL inside its encoder, so threads scale like the reference's own 5-thread pool."""
    from concurrent.futures import ThreadPoolExecutor
    from oracle.pillow_path import prepare_page_cpu

    def one(im):
        png, b64, _ = prepare_page_cpu(im, **kw)
        return len(png)
    best = None
    with ThreadPoolExecutor(threads) as ex:
        for _ in range(repeats):
            t0 = time.perf_counter()
            sizes = list(ex.map(one, pages))
            dt = time.perf_counter() - t0
            best = dt if best is None else min(best, dt)
    return len(pages) / best, sizes


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    cores = host_cores()
    fac = PageFactory(cores)
    pages = [to_pil(a) for a in fac.arrays(c2_specs(0, PAGES_PER_GPU))]
    fac.close()
    for _ in range(max(0, min(args.warmup, 1))):
        cpu_path_pages_per_s(pages[:max(1, cores // 2)], cores)
    t0 = time.perf_counter()
    n = 0
    for _ in range(args.steps):
        cpu_path_pages_per_s(pages, cores)
        n += len(pages)
    dt = time.perf_counter() - t0
    v = n / dt
    line = {
        "impl": "reference", "metric": "pages/sec (resize+PNG+base64)", "value": v, "unit": "pages/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * dt / max(1, args.steps), "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "u8", "data": "synthetic",
        "config": bench_config(),
        "cpu_baseline": {"value": v, "unit": "pages/s", "cores": cores, "kind": "reference",
                         "sample": f"all {len(pages)} pages of the step, Pillow {__import__('PIL').__version__} Image.save(PNG)+base64 on {cores} threads"},
        "e2e": {"value": v, "unit": "pages/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line), flush=True)
    return 0


# ------------------------------------------------------------------------------------------------ clocks
class ClockSampler:
    Q = "clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown," \
        "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"

    def __init__(self, index: int):
        self.rows, self.proc = [], None
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "100",
                                          "-i", str(index)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for ln in self.proc.stdout:
            self.rows.append((time.perf_counter(), ln.strip()))

    def stop(self, t0: float, t1: float):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        sm, mx, reasons = [], None, set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        rows = [r for t, r in self.rows if t0 <= t <= t1] or [r for _, r in self.rows[-3:]]
        for r in rows:
            f = [x.strip() for x in r.split(",")]
            try:
                sm.append(float(f[0])); mx = float(f[1])
            except (ValueError, IndexError):
                continue
            for nm, val in zip(names, f[2:6]):
                if val.lower().startswith("active"):
                    reasons.add(nm)
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": mx, "reasons": sorted(reasons), "samples": len(sm)}


# ------------------------------------------------------------------------------------------------ GPU arm helpers
def pinned_like(torch, arrays):
    """One pinned host block holding copies of `arrays` (same shapes); returns the list of numpy views, or None if pinning fails."""
    import numpy as np
    total = sum(int(a.nbytes) for a in arrays)
    try:
        block = torch.empty(total, dtype=torch.uint8, pin_memory=True)
    except RuntimeError:
        return None
    out, off = [], 0
    flat = block.numpy()
    for a in arrays:
        v = flat[off:off + a.nbytes].reshape(a.shape)
        np.copyto(v, a)
        out.append(v)
        off += a.nbytes
    out.append(block)                                   # keep-alive rides at the end of the list
    return out


def device_resident_rate(eng, torch, N, arrays, kw, steps, warm=2):
    """pages/s of one vcp_prepare_batch over `arrays` already in HBM, outputs left in HBM; also the per-stage CUDA-event times."""
    from vision_compression_project_b200.api import PagePrep, _as_source
    dev = [torch.from_numpy(a).cuda() for a in arrays]
    n = len(dev)
    descs = (N.PageDesc * n)()
    for i, t in enumerate(dev):
        descs[i] = PagePrep._plan(_as_source(t, None), None, kw.get("max_side"), "RGB", 1, kw.get("reducing_gap"))
    o = N.Opts()
    o.out_channels, o.resample, o.compress_level, o.want_b64, o.src_device, o.dst_device = 3, 1, 6, 1, 1, 1
    bp, bb = eng.output_bound(descs, n, o)
    cap_p, cap_b = max(64 << 20, bp // 3), max(88 << 20, bb // 3)
    op = torch.empty(cap_p, dtype=torch.uint8, device="cuda")
    ob = torch.empty(cap_b, dtype=torch.uint8, device="cuda")
    for _ in range(warm):
        eng.run(descs, n, o, op.data_ptr(), cap_p, ob.data_ptr(), cap_b)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    stage, launches = {}, 0
    torch.cuda.synchronize(); e0.record()
    for _ in range(steps):
        eng.run(descs, n, o, op.data_ptr(), cap_p, ob.data_ptr(), cap_b)
        st = eng.stats()
        launches += st["kernel_launches"]
        for k_, v_ in st.items():
            if k_.startswith("ms_"):
                stage[k_] = stage.get(k_, 0.0) + v_ / steps
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / steps
    del dev, op, ob
    return n / (ms / 1e3), ms, {k_: round(v_, 4) for k_, v_ in stage.items()}, launches


def check_against_pillow(outs, arrays, idx, kw):
    """Outside every timed region: pages `idx` decode to exactly the pixels of the Pillow path; returns ours/Pillow PNG size over them."""
    import io
    from PIL import Image
    from oracle.pillow_path import prepare_page_cpu
    ours = ref = 0
    for i in idx:
        png, _, exp = prepare_page_cpu(to_pil(arrays[i]), **kw)
        dec = Image.open(io.BytesIO(outs[i].png)); dec.load()
        assert dec.size == exp.size and dec.mode == exp.mode and dec.tobytes() == exp.tobytes(), f"page {i}: pixels differ from the Pillow path"
        assert outs[i].b64 == base64.b64encode(outs[i].png), f"page {i}: base64 mismatch"
        ours += len(outs[i].png); ref += len(png)
    return ours / max(1, ref)


def sub_config(name, fn):
    """A side config must never take the headline down with it: its failure is reported in its record."""
    t0 = time.perf_counter()
    try:
        rec = fn()
    except Exception as e:  # noqa: BLE001
        rec = {"error": f"{type(e).__name__}: {e}"[:300]}
    rec["bench_wall_s"] = round(time.perf_counter() - t0, 2)
    return rec


def lz_traffic():
    """DRAM bytes of one k_lz launch on the C2 batch from the committed ncu capture — only if it was taken from this kernel source
    (the file records the sha256 of deflate_lz.cu; .git does not travel to the GPU box, so HEAD cannot be compared there)."""
    try:
        cands = sorted(f for f in os.listdir(os.path.join(ROOT, "profiles")) if f.endswith("_k_lz_traffic.json"))
        tj = json.load(open(os.path.join(ROOT, "profiles", cands[-1])))
        src = open(os.path.join(ROOT, "vision_compression_project_b200", "csrc", "deflate_lz.cu"), "rb").read()
        if tj.get("source_sha16") != hashlib.sha256(src).hexdigest()[:16]:
            return None, f"{cands[-1]} is from another version of deflate_lz.cu"
        return tj["dram_bytes_read"] + tj["dram_bytes_write"], cands[-1]
    except Exception as e:  # noqa: BLE001
        return None, f"no capture ({type(e).__name__})"


# ------------------------------------------------------------------------------------------------ GPU arm
def run_ours(args):
    import numpy as np
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    K, W = args.steps, max(args.warmup, 3)
    cores = host_cores()
    from vision_compression_project_b200 import sharding

    import torch
    import torch.distributed as dist
    from vision_compression_project_b200 import _native as N
    from vision_compression_project_b200.api import PagePrep
    import vision_compression_project_b200 as V
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device — this path has no CPU fallback (use --impl reference for the CPU arm)")
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    all_cores = sorted(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else []
    my_cores = sharding.pin_rank_to_cores(local, int(os.environ.get("LOCAL_WORLD_SIZE", world))) if not args.no_pin else []
    kept_heap = sharding.keep_result_memory()              # results (Python bytes) reuse heap pages instead of faulting in fresh ones
    cores = host_cores() if world > 1 and not args.no_pin else cores      # of this rank from here on

    # ---- synthetic pages: drawn by spawned worker processes (FreeType rendering holds the GIL; spawn, not fork, so CUDA in this
    #      process is no concern), written straight into pinned host memory
    t_gen = time.perf_counter()
    fac = PageFactory(max(2, cores if my_cores and world > 1 else cores // world))
    n = PAGES_PER_GPU
    w, h = 1700, 2200
    host = torch.empty((n, h, w, 3), dtype=torch.uint8, pin_memory=True)          # pinned host copies of the pages
    arrs = fac.arrays(c2_specs(rank * PAGES_PER_GPU, PAGES_PER_GPU), out=[host[i].numpy() for i in range(n)])
    c4_lo, c4_hi = sharding.page_range(args.c4_pages, rank, world)
    c4_arrs, c4_pinned = [], False
    if args.c4_pages:
        m4 = c4_hi - c4_lo
        try:
            c4_block = torch.empty((m4, h, w, 3), dtype=torch.uint8, pin_memory=True)
            c4_pinned = True
        except RuntimeError:                                                       # not enough lockable memory on this box: pageable arrays
            c4_block = torch.empty((m4, h, w, 3), dtype=torch.uint8)
        c4_arrs = fac.arrays(c4_specs(c4_lo, c4_hi), out=[c4_block[i].numpy() for i in range(m4)])
    side = {}
    if world == 1 and not args.headline_only:
        side["S150"] = fac.arrays([(s_, "letter", 150, "RGB", s_ % 4 == 3) for s_ in range(64)])
        side["C3"] = fac.arrays(c3_specs(args.c3_pages))
        side["C5"] = fac.arrays(c5_specs())
    fac.close()
    t_gen = time.perf_counter() - t_gen

    dev = host.cuda()
    eng = PagePrep(local)

    descs = (N.PageDesc * n)()
    for i in range(n):
        descs[i].src, descs[i].width, descs[i].height, descs[i].channels = dev[i].data_ptr(), w, h, 3
    opts = N.Opts()
    opts.out_channels, opts.resample, opts.compress_level, opts.want_b64, opts.src_device, opts.dst_device = 3, 1, 6, 1, 1, 1
    bp, bb = eng.output_bound(descs, n, opts)
    cap_p, cap_b = max(64 << 20, bp // 3), max(88 << 20, bb // 3)
    out_p = torch.empty(cap_p, dtype=torch.uint8, device="cuda")
    out_b = torch.empty(cap_b, dtype=torch.uint8, device="cuda")

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(x: float) -> float:
        t = torch.tensor([x], device="cuda", dtype=torch.float64)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    def step_device():
        return eng.run(descs, n, opts, out_p.data_ptr(), cap_p, out_b.data_ptr(), cap_b)

    # ---- device-resident: value + per-kernel times (CUDA events on the library's own stream, read from vcp_stats)
    for _ in range(W):
        res = step_device()
    png_bytes = sum(r.png_len for r in res); b64_bytes = sum(r.b64_len for r in res)
    in_bytes = n * h * w * 3
    sampler = ClockSampler(local) if rank == 0 else None
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    stage = {}
    launches = 0
    t0 = time.perf_counter(); e0.record()
    for _ in range(K):
        step_device()
        st = eng.stats()
        launches += st["kernel_launches"]
        for k_, v_ in st.items():
            if k_.startswith("ms_"):
                stage[k_] = stage.get(k_, 0.0) + v_
    e1.record(); barrier()
    ms_step = max_over_ranks(e0.elapsed_time(e1)) / K
    value = world * n / (ms_step / 1e3)

    # ---- end to end through the public function: pinned host pages in, Python bytes out
    host_np = [host[i].numpy() for i in range(n)]
    for _ in range(2):
        outs = eng.prepare_pages(host_np)
    assert all(o.error is None for o in outs)
    barrier()
    Ke = max(1, min(K, 10))
    t0e = time.perf_counter()
    for _ in range(Ke):
        outs = eng.prepare_pages(host_np)
    torch.cuda.synchronize()
    e2e_sync = world * n * Ke / max_over_ranks(time.perf_counter() - t0e)       # one synchronous prepare_pages call per step
    # the same K steps through prepare_stream: several steps in flight, so one step's pipeline drains (LZ / Huffman / D2H of its
    # last pages with the PCIe link idle) while the next one copies.  Every step still moves its 718 MB in and its bytes out.
    depth = 2                                                # batches in flight: measured 2 >= 3 > 4 on B200 (tools/quick_e2e.py): more only fight for the one PCIe link
    for _o in V.prepare_stream((host_np for _ in range(4)), depth=depth, device=local):
        assert all(o.error is None for o in _o)
    barrier()
    t0e = time.perf_counter()
    n_done = 0
    for _o in V.prepare_stream((host_np for _ in range(Ke)), depth=depth, device=local):
        n_done += len(_o)
    torch.cuda.synchronize()
    dte = time.perf_counter() - t0e
    assert n_done == n * Ke and _o[0].png == outs[0].png and _o[-1].b64 == outs[-1].b64
    e2e_v = world * n * Ke / max_over_ranks(dte)
    clocks = sampler.stop(t0, time.perf_counter()) if sampler else None      # both timed regions (device-resident steps and e2e steps)
    e2e_detail = dict(getattr(eng, "last_timing", {}))
    e2e_detail.update({k_: v_ for k_, v_ in eng.stats().items() if k_.startswith("ms_")})
    # the same call with PIL images in (the reference's own input type): Pillow's pixel storage is read in place (pageable memory)
    pages = [to_pil(a) for a in arrs]
    for _ in range(2):
        outs_pil = eng.prepare_pages(pages)
    barrier()
    t0p = time.perf_counter()
    Kp = max(1, min(K, 5))
    for _ in range(Kp):
        outs_pil = eng.prepare_pages(pages)
    e2e_pil = world * n * Kp / max_over_ranks(time.perf_counter() - t0p)
    assert outs_pil[0].png == outs[0].png and outs_pil[-1].png == outs[-1].png
    # the reference's calling pattern (pdf_extract.py:313-333): 5 worker threads, each page its own prepare_page(PIL image) call
    def five_threads(reps):
        errs = []

        def work(k):
            try:
                for r_ in range(reps):
                    for i in range(k, n, 5):
                        V.prepare_page(pages[i], device=local)
            except Exception as e:  # noqa: BLE001
                errs.append(e)
        ths = [threading.Thread(target=work, args=(k,)) for k in range(5)]
        t_ = time.perf_counter()
        [x.start() for x in ths]; [x.join() for x in ths]
        if errs:
            raise errs[0]
        return reps * n / (time.perf_counter() - t_)
    five_threads(1)
    barrier()
    e2e_five = five_threads(2)
    t_five = torch.tensor([e2e_five], device="cuda", dtype=torch.float64)
    if world > 1:
        dist.all_reduce(t_five, op=dist.ReduceOp.SUM)                        # independent ranks: rates add
    e2e_five = float(t_five.item())

    # ---- C4: the 2,000-page document, STRONG scaling — page ranges from sharding.page_range, every rank streams its range through
    #      prepare_stream (host arrays in, Python bytes out, batches of 64 pages), no exchange between ranks while the document is
    #      processed; afterwards rank 0 collects the bytes in page order (gather_in_page_order: control plane, timed separately)
    c4 = None
    if args.c4_pages:
        def run_c4():
            src = c4_arrs
            batches = [src[i:i + 64] for i in range(0, len(src), 64)]
            for _o in V.prepare_stream(iter(batches[:2]), depth=depth, device=local):
                pass
            best, per_rank, outs4 = None, None, None
            n_mine = c4_hi - c4_lo
            for _rep in range(2):
                # a rank consumes its pages batch by batch, as the reference's workers write their page file and drop the image
                # (pdf_extract.py:130): every result is produced as Python bytes and looked at, the first and last page are kept
                # for the pixel check; nothing else stays alive, so result memory is reused instead of faulted in 1.7 MB per page
                barrier()
                t_ = time.perf_counter()
                n_done, png_sum, b64_sum, first, last_ = 0, 0, 0, None, None
                for b_ in V.prepare_stream(iter(batches), depth=depth, device=local):
                    for o in b_:
                        assert o.error is None
                        png_sum += len(o.png); b64_sum += len(o.b64)
                        if n_done == 0: first = o
                        last_ = o
                        n_done += 1
                dt_ = time.perf_counter() - t_
                assert n_done == n_mine
                tt = torch.zeros(world, device="cuda", dtype=torch.float64); tt[rank] = dt_
                if world > 1:
                    dist.all_reduce(tt)
                tmax = float(tt.max().item())
                if best is None or tmax < best:
                    best, per_rank = tmax, [round(1e3 * float(x), 1) for x in tt.tolist()]
            rec = {"pages": args.c4_pages, "scaling": "strong", "pages_per_s": args.c4_pages / best, "per_rank_ms": per_rank,
                   "page_ranges": [list(sharding.page_range(args.c4_pages, r_, world)) for r_ in range(world)],
                   "host_memory": "pinned" if c4_pinned else "pageable", "api": f"sharding.page_range + prepare_stream(batches of 64, depth={depth}) per rank",
                   "what": "2,000 letter-200 pages, every 4th photo-heavy, host arrays in -> PNG + base64 bytes on the rank that owns the page "
                           "(the reference's workers each write their own page file, pdf_extract.py:130), results consumed batch by batch; "
                           "fastest of 2 passes, max over ranks"}
            if world > 1:
                outs4 = [o for b_ in V.prepare_stream(iter(batches), depth=depth, device=local) for o in b_]      # untimed pass that keeps every result
                barrier()
                t_ = time.perf_counter()
                allp = sharding.gather_in_page_order([(o.png, o.b64) for o in outs4], c4_lo, args.c4_pages)
                rec["gather_to_rank0_ms"] = round(1e3 * max_over_ranks(time.perf_counter() - t_), 1)
                if rank == 0:
                    assert len(allp) == args.c4_pages and all(p is not None for p in allp) and allp[c4_lo][0] == outs4[0].png
                    rec["gathered_pages_per_s"] = args.c4_pages / (best + rec["gather_to_rank0_ms"] / 1e3)
            rec["png_size_vs_pillow"] = check_against_pillow([first, last_], [c4_arrs[0], c4_arrs[-1]], [0, 1], {}) if rank == 0 else None
            rec["png_bytes_per_page"] = png_sum / max(1, n_mine)
            return rec
        c4 = sub_config("C4", run_c4)

    if rank == 0:
        if all_cores and hasattr(os, "sched_setaffinity"):           # the CPU legs below use the whole box again (the other ranks are idle)
            os.sched_setaffinity(0, all_cores)
            cores = len(all_cores)
        # correctness of what was timed (not timed): first page decodes to the input and base64 matches
        import io
        from PIL import Image
        dec = Image.open(io.BytesIO(outs[0].png)); dec.load()
        assert dec.tobytes() == pages[0].tobytes() and outs[0].b64 == base64.b64encode(outs[0].png)
        # the reverse path, reported beside the headline: this step's PNGs (host bytes) -> pixels in HBM on the GPU,
        # and Pillow's decoder on the host threads for the same files
        pngs = [o.png for o in outs]
        back = eng.decode_pages(pngs, to_device=True)
        assert all(not isinstance(b_, Exception) for b_ in back) and bytes(back[1].cpu().numpy().tobytes()) == pages[1].tobytes()
        del back
        torch.cuda.synchronize(); t0d = time.perf_counter()
        for _ in range(3):
            back = eng.decode_pages(pngs, to_device=True); del back
        torch.cuda.synchronize()
        dec_v = 3 * n / (time.perf_counter() - t0d)
        from concurrent.futures import ThreadPoolExecutor

        def _pil_dec(b_):
            im_ = Image.open(io.BytesIO(b_)); im_.load(); return im_.size
        with ThreadPoolExecutor(cores) as ex:
            list(ex.map(_pil_dec, pngs[:cores]))
            t0d = time.perf_counter(); list(ex.map(_pil_dec, pngs)); dec_cpu = len(pngs) / (time.perf_counter() - t0d)
        decode_info = {"value": dec_v, "unit": "pages/s", "batch": n, "what": "vcp_png_decode_batch: this step's PNG bytes (host) -> pixels in HBM, pixel-checked, Adler-32 verified",
                       "pillow_cpu_pages_per_s": dec_cpu, "cores": cores}
        peaks = {}
        try:
            peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        except Exception:
            pass
        peak = float(peaks.get("hbm_gbs", 6650.0))
        alg_bytes = in_bytes + png_bytes + b64_bytes                       # SURVEY.md §8 d: B_in + B_png + B_b64 per page, x pages
        per = {k_: v_ / K for k_, v_ in stage.items()}
        dom = max((k_ for k_ in per if k_ not in ("ms_total", "ms_h2d", "ms_d2h")), key=lambda k_: per[k_])
        dom_ms = per[dom]
        achieved = alg_bytes / (dom_ms / 1e3) / 1e9
        traffic, traffic_src = lz_traffic() if dom == "ms_lz" else (None, "dominant kernel is not k_lz")
        cpu_v, cpu_sizes = cpu_path_pages_per_s(pages, cores, repeats=2)
        ratio = sum(len(o.png) for o in outs) / max(1, sum(cpu_sizes))

        # ---- the other BASELINE configs (N = 1 runs): device-resident rate, end-to-end rate, size against Pillow, pixel check
        configs = {}
        if world == 1 and not args.headline_only:
            def run_c1():
                ref = Image.open(os.path.join(ROOT, "tests", "golden", "ref_page_1.png")); ref.load()
                a = np.ascontiguousarray(np.asarray(ref))
                for _ in range(3):
                    r1 = V.prepare_page(ref)
                t_ = time.perf_counter()
                for _ in range(20):
                    r1 = V.prepare_page(ref)
                lat_pil = (time.perf_counter() - t_) / 20
                t_ = time.perf_counter()
                for _ in range(20):
                    V.prepare_page(a)
                lat_np = (time.perf_counter() - t_) / 20
                dev_rate, dev_ms, st_, _ = device_resident_rate(eng, torch, N, [a], {}, 20)
                t_ = time.perf_counter()
                from oracle.pillow_path import prepare_page_cpu
                for _ in range(3):
                    prepare_page_cpu(ref)
                cpu_ms = (time.perf_counter() - t_) / 3 * 1e3
                return {"workload": "C1: the reference's recorded output/page_1.png (1654x2339 RGB), one prepare_page call", "pages": 1,
                        "e2e_pages_per_s": 1 / lat_pil, "latency_ms_pil_in": 1e3 * lat_pil, "latency_ms_numpy_in": 1e3 * lat_np,
                        "device_resident_pages_per_s": dev_rate, "device_ms": dev_ms, "stage_ms": st_,
                        "pillow_cpu_ms": cpu_ms, "png_bytes": len(r1.png),
                        "png_size_vs_pillow": check_against_pillow([r1], [a], [0], {})}

            def run_side(label, arrays, kw, workload, cpu_sample):
                dev_rate, dev_ms, st_, _ = device_resident_rate(eng, torch, N, arrays, kw, 3)
                pin = pinned_like(torch, arrays)
                src = pin[:-1] if pin else arrays
                outs_ = eng.prepare_pages(src, **kw)
                t_ = time.perf_counter()
                for _ in range(2):
                    outs_ = eng.prepare_pages(src, **kw)
                e2e_ = 2 * len(src) / (time.perf_counter() - t_)
                assert all(o.error is None for o in outs_)
                pil_ = [to_pil(a) for a in arrays]
                eng.prepare_pages(pil_, **kw)                        # warm: the pinned bounce buffers grow to this batch's group sizes once
                t_ = time.perf_counter()
                outs_p = eng.prepare_pages(pil_, **kw)
                e2e_p = len(pil_) / (time.perf_counter() - t_)
                assert outs_p[0].png == outs_[0].png and outs_p[-1].png == outs_[-1].png
                idx = sorted(set([0, len(arrays) - 1] + list(range(0, len(arrays), max(1, len(arrays) // cpu_sample)))))
                ratio_ = check_against_pillow(outs_, arrays, idx, kw)
                cpu_rate, _ = cpu_path_pages_per_s([pil_[i] for i in idx], cores, **kw)
                return {"workload": workload, "pages": len(arrays), "device_resident_pages_per_s": dev_rate, "device_ms": dev_ms, "stage_ms": st_,
                        "e2e_pages_per_s": e2e_, "e2e_pil_images_in_pages_per_s": e2e_p, "host_memory": "pinned" if pin else "pageable",
                        "sizes_out": sorted({o.size for o in outs_})[:4], "png_bytes_per_page": sum(len(o.png) for o in outs_) / len(outs_),
                        "png_size_vs_pillow": ratio_, "pixel_checked_pages": len(idx),
                        "pillow_cpu_pages_per_s": cpu_rate, "cpu_sample": f"{len(idx)} of the pages on {cores} threads"}
            configs["C1"] = sub_config("C1", run_c1)
            configs["S150"] = sub_config("S150", lambda: run_side(
                "S150", side["S150"], {},
                "S150: the service's default rasterisation (backend/app/config.py:57 DEFAULT_DPI = 150): 64 letter pages @150 DPI (1275x1650 RGB, every 4th photo-heavy), convert('RGB') + PNG + base64", 16))
            configs["C3"] = sub_config("C3", lambda: run_side(
                "C3", side["C3"], {"max_side": MAX_SIDE},
                f"C3: {len(side['C3'])} synthetic letter pages @300 DPI (2550x3300 RGB, every 4th photo-heavy) -> LANCZOS to {MAX_SIDE} px long edge (1212x1568) + PNG + base64", 16))
            configs["C5"] = sub_config("C5", lambda: run_side(
                "C5", side["C5"], {"max_side": MAX_SIDE, "reducing_gap": 2.0},
                f"C5: 48 page types (A4/letter/legal x 150/200/300/600 DPI x L/RGB x text/photo), thumbnail rule to {MAX_SIDE} px with reducing_gap=2.0, convert('RGB') + PNG + base64", 48))
        line = {
            "metric": "pages/sec (resize+PNG+base64)", "value": value, "unit": "pages/s", "n_gpus": world, "steps": K, "warmup": W,
            "ms_per_step": ms_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "u8",
            "data": "synthetic",
            "config": bench_config(),
            "size": {"png_bytes_per_page": png_bytes / n, "png_size_vs_pillow": ratio, "tolerance": 1.05},
            "e2e": {"value": e2e_v, "unit": "pages/s", "h2d_bytes_per_step": in_bytes, "d2h_bytes_per_step": png_bytes + b64_bytes,
                    "api": f"prepare_stream(K batches of pinned uint8 arrays, depth={depth}) -> PreparedPage(png bytes, b64 bytes) per page",
                    "single_call_pages_per_s": e2e_sync,
                    "last_step_ms": {k_: round(v_, 3) for k_, v_ in e2e_detail.items()},
                    "pil_images_in_pages_per_s": e2e_pil,
                    "five_thread_single_page_pages_per_s": e2e_five},
            "gpu_launches": int(launches),
            "clocks": clocks,
            "roofline": {"bound": "hbm", "kernel": {"ms_lz": "k_lz", "ms_filter": "k_png_filter", "ms_huff": "k_huff_build/k_layout/k_payload_init/k_huff_emit",
                                                    "ms_b64": "k_base64_pages", "ms_assemble": "k_png_finish", "ms_convert": "pixel kernels",
                                                    "ms_resample": "k_resample_h/k_resample_v"}.get(dom, dom),
                         "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                         "peak_source": "MEASURED_PEAKS.json hbm_gbs (of measured)" if peaks else "fallback 6650 GB/s (of fallback)",
                         "traffic": traffic, "traffic_source": traffic_src, "algorithmic_bytes_per_launch": alg_bytes,
                         "kernel_own_bytes": int(1.066 * n * h * (1 + 3 * w)),   # k_lz itself: reads the filtered stream once, writes ~0.07 B of tokens per byte
                         "kernel_ms": dom_ms,
                         "stage_ms": per},
            "cpu_baseline": {"value": cpu_v, "unit": "pages/s", "cores": cores, "kind": "reference",
                             "sample": f"all {n} pages of the batch, best of 2, Pillow Image.save(PNG)+base64 on {cores} threads"},
            "decode": decode_info,
            "configs": configs,
            "c4": c4,
            "host": {"cores": len(all_cores) or cores, "cores_per_rank": len(my_cores) or cores, "pinned_ranks": bool(my_cores) and world > 1, "allocator": "glibc mallopt: freed result blocks stay in the heap (sharding.keep_result_memory)" if kept_heap else "default"},
            "page_generation_s": round(t_gen, 1),
        }
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    return 0


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--c3-pages", type=int, default=C3_PAGES, help="pages of the C3 side config (BASELINE: 256)")
    ap.add_argument("--c4-pages", type=int, default=C4_PAGES, help="pages of the C4 document (BASELINE: 2000; 0 = skip the leg)")
    ap.add_argument("--no-pin", action="store_true", help="N > 1: leave the ranks' threads unpinned (default: each rank gets its slice of the cores)")
    ap.add_argument("--headline-only", action="store_true", help="C2 only (profiling runs): no C1/C3/C5 side configs")
    args = ap.parse_args()
    if args.impl == "reference":
        return run_reference(args)
    return run_ours(args)


if __name__ == "__main__":
    sys.exit(main())
