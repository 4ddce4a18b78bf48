"""Page-range sharding across the GPUs of one box (SURVEY.md §8 e): pages are independent, so every rank owns a
contiguous range and there is no exchange step (no collective on the data path).

The reference's analogue is the 5-thread page pool of `extract_pdf_to_page_jsons`
(backend/app/pipeline/pdf_extract.py:313-350): results are collected in page order, failed pages are reported, not fatal.
"""
from __future__ import annotations

from typing import List, Sequence, Tuple


def page_range(n_pages: int, rank: int, world: int) -> Tuple[int, int]:
    """Contiguous [lo, hi) of rank `rank`; ranges tile [0, n_pages) and differ in size by at most one page."""
    if world <= 0 or not 0 <= rank < world:
        raise ValueError("bad rank/world")
    base, extra = divmod(n_pages, world)
    lo = rank * base + min(rank, extra)
    return lo, lo + base + (1 if rank < extra else 0)


def balanced_ranges(weights: Sequence[int], world: int) -> List[Tuple[int, int]]:
    """Contiguous ranges with near-equal total weight (weight = W*H*C of a page) for mixed-size batches (config C5).
    Greedy prefix cut at k/world of the total; every rank gets a (possibly empty) range, order is preserved."""
    if world <= 0:
        raise ValueError("bad world")
    total = sum(weights)
    out, lo, acc = [], 0, 0
    n = len(weights)
    for r in range(world):
        target = total * (r + 1) / world
        hi = lo
        while hi < n and (acc + weights[hi] / 2 <= target or r == world - 1):
            acc += weights[hi]
            hi += 1
        out.append((lo, hi))
        lo = hi
    return out


def prepare_pages_sharded(images: Sequence, rank: int, world: int, device: int | None = None, **kw):
    """This rank's share of `images` through the GPU path. Returns (lo, results) so the caller can place them in page order."""
    from .api import prepare_pages
    lo, hi = page_range(len(images), rank, world)
    return lo, prepare_pages(images[lo:hi], device=rank if device is None else device, **kw)


def prepare_pages_all_gpus(images: Sequence, devices: Sequence[int] | None = None, **kw) -> list:
    """One process, every GPU of the box: contiguous page ranges (balanced by W*H*C), one host thread per device (the C ABI
    releases the GIL), results back in page order.  No exchange between devices — the multi-GPU analogue of the reference's
    page pool (pdf_extract.py:313-350)."""
    import threading
    import torch
    from .api import _as_source, prepare_pages
    if devices is None:
        devices = list(range(torch.cuda.device_count()))
    if not devices:
        raise RuntimeError("no CUDA device")
    if len(devices) == 1 or len(images) <= 1:
        return prepare_pages(images, device=devices[0], **kw)

    def weight(im):
        try:
            s = _as_source(im, kw.get("raw_shape"))
            return s.w * s.h * s.c
        except (ValueError, TypeError):
            return 1
    ranges = balanced_ranges([weight(im) for im in images], len(devices))
    out: list = [None] * len(images)
    errs: list = []

    def work(dev, lo, hi):
        try:
            if hi > lo:
                out[lo:hi] = prepare_pages(images[lo:hi], device=dev, **kw)
        except BaseException as e:  # noqa: BLE001 - re-raised in the caller
            errs.append(e)
    ths = [threading.Thread(target=work, args=(d, lo, hi)) for d, (lo, hi) in zip(devices, ranges)]
    [t.start() for t in ths]
    [t.join() for t in ths]
    if errs:
        raise errs[0]
    return out


def _flatten_bytes(items):
    """items: each a bytes object or a tuple/list of bytes-or-None.  Returns (layout, parts) or None when something else is inside."""
    layout, parts = [], []
    for it in items:
        if isinstance(it, (bytes, bytearray, memoryview)):
            layout.append(len(it)); parts.append(it)
        elif isinstance(it, (tuple, list)) and all(x is None or isinstance(x, (bytes, bytearray, memoryview)) for x in it):
            layout.append(tuple(-1 if x is None else len(x) for x in it))
            parts.extend(x for x in it if x is not None)
        else:
            return None
    return layout, parts


def gather_in_page_order(local: Sequence, lo: int, n_pages: int, group=None) -> list | None:
    """Host-side concatenation of per-rank results in page order on rank 0 (variable-length byte strings; this is
    control-plane traffic over the default process group, not a data-path collective).

    One process per GPU means all ranks share the host: byte payloads (PNG / base64 strings, or tuples of them) travel through
    one POSIX shared-memory block per rank — a memcpy in, a memcpy out — and only the layout goes through the process group
    (pickling 1.4 GB of page bytes through gather_object took 8.7 s for the 2,000-page document on 2 ranks; this takes well under
    a second).  Anything else, or ranks on different hosts, falls back to gather_object."""
    import os
    import torch.distributed as dist
    world = dist.get_world_size(group)
    rank = dist.get_rank(group)
    same_host = int(os.environ.get("LOCAL_WORLD_SIZE", "0") or 0) == world or world == 1
    flat = _flatten_bytes(local) if same_host else None
    if flat is not None:                               # a tmpfs too small for the block would fault on write, not raise: check first
        try:
            vfs = os.statvfs("/dev/shm")
            if vfs.f_bavail * vfs.f_frsize < 2 * sum(len(p) for p in flat[1]) * world + (64 << 20):
                flat = None
        except OSError:
            flat = None
    ok = [None] * world
    dist.all_gather_object(ok, flat is not None, group=group)
    if not all(ok):
        gathered = [None] * world if rank == 0 else None
        dist.gather_object((lo, list(local)), gathered, dst=0, group=group)
        if rank != 0:
            return None
        out = [None] * n_pages
        for l, items in gathered:
            out[l:l + len(items)] = items
        return out
    import ctypes
    from multiprocessing import shared_memory
    from . import _native
    threads = max(2, min(16, (os.cpu_count() or 4) // 2))
    layout, parts = flat
    total = sum(len(p) for p in parts)
    shm = shared_memory.SharedMemory(create=True, size=max(1, total))
    try:
        off = 0
        for p in parts:
            shm.buf[off:off + len(p)] = p
            off += len(p)
        metas = [None] * world if rank == 0 else None
        dist.gather_object((lo, shm.name, layout), metas, dst=0, group=group)
        out = None
        if rank == 0:
            out = [None] * n_pages
            for l, name, lay in metas:
                blk = shm if name == shm.name else shared_memory.SharedMemory(name=name)
                try:
                    # every byte string of the block in one multi-threaded copy (vcp_host_scatter fills fresh bytes objects, GIL released)
                    ranges, o = [], 0
                    for ent in lay:
                        for ln in ((ent,) if isinstance(ent, int) else ent):
                            if ln >= 0:
                                ranges.append((o, ln)); o += ln
                    anchor = ctypes.c_char.from_buffer(blk.buf)
                    try:
                        objs = iter(_native.gather_bytes(ctypes.addressof(anchor), ranges, threads))
                    finally:
                        del anchor
                    for k, ent in enumerate(lay):
                        out[l + k] = next(objs) if isinstance(ent, int) else tuple(None if ln < 0 else next(objs) for ln in ent)
                finally:
                    if blk is not shm:
                        blk.close()
        dist.barrier(group=group)                   # rank 0 has read every block
        return out
    finally:
        shm.close()
        shm.unlink()


def pin_rank_to_cores(rank: int, world: int) -> list:
    """One process per GPU on a shared host: give rank r its own slice of the cores this process may use, so that the ranks' pack /
    scatter threads (and the Python threads that feed them) do not migrate over each other.  Returns the cores of this rank
    (no-op outside Linux).  The library sizes its host thread pools from LOCAL_WORLD_SIZE; this only decides WHERE they run."""
    import os
    if not hasattr(os, "sched_setaffinity") or world <= 1:
        return sorted(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else []
    cores = sorted(os.sched_getaffinity(0))
    lo, hi = page_range(len(cores), rank, world)
    mine = cores[lo:hi] or cores
    os.sched_setaffinity(0, mine)
    return mine


def keep_result_memory(mmap_threshold: int = 1 << 30, trim_threshold: int = (1 << 31) - 1, top_pad: int = 1 << 28) -> bool:
    """A rank turns its pages into Python `bytes` — megabytes each, gigabytes per second.  glibc serves such blocks with mmap and gives
    them back with munmap, so every result is written into pages the kernel has just zeroed and mapped: with few host cores per
    rank (4 at 8 GPUs on a 32-core box) that page-fault work, not the GPU or the link, sets the rate (one rank on 4 cores: 3 170 ->
    4 260 pages/s on the C4 document with these settings).  This tells the allocator of THIS process to keep freed blocks in its
    heap (glibc `mallopt`: M_MMAP_THRESHOLD, M_TRIM_THRESHOLD, M_TOP_PAD); the price is that the heap keeps the high-water mark of
    the results that were alive at once.  Returns False where there is no glibc `mallopt`."""
    import ctypes
    try:
        libc = ctypes.CDLL("libc.so.6")
        ok = libc.mallopt(-3, int(mmap_threshold)) and libc.mallopt(-1, int(trim_threshold)) and libc.mallopt(-2, int(top_pad))
        return bool(ok)
    except (OSError, AttributeError):
        return False
