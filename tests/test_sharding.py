"""CPU: page-range sharding, including a world-size-2 gloo run of the gather (the N>1 host path; no GPU)."""
import os
import socket

import pytest
import torch.distributed as dist
import torch.multiprocessing as mp

from vision_compression_project_b200 import sharding as S


def test_page_range_tiles():
    for n in (0, 1, 7, 64, 2000, 2001):
        for world in (1, 2, 4, 8):
            rs = [S.page_range(n, r, world) for r in range(world)]
            assert rs[0][0] == 0 and rs[-1][1] == n
            assert all(a[1] == b[0] for a, b in zip(rs, rs[1:]))
            sizes = [hi - lo for lo, hi in rs]
            assert max(sizes) - min(sizes) <= 1
    with pytest.raises(ValueError):
        S.page_range(4, 2, 2)


def test_balanced_ranges_mixed_sizes():
    w = [1240 * 1754, 5100 * 8400 * 3, 1700 * 2200 * 3, 2550 * 3300, 4960 * 7016 * 3, 1275 * 1650 * 3] * 5
    for world in (1, 2, 4, 8):
        rs = S.balanced_ranges(w, world)
        assert len(rs) == world and rs[0][0] == 0 and rs[-1][1] == len(w)
        assert all(a[1] == b[0] for a, b in zip(rs, rs[1:]))
        loads = [sum(w[lo:hi]) for lo, hi in rs]
        assert max(loads) <= sum(w) / world + max(w)


def _worker(rank, world, port, n_pages, q, same_host=False):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    if same_host:
        os.environ["LOCAL_WORLD_SIZE"] = str(world)                          # as torchrun sets it: payloads go through shared memory
    dist.init_process_group("gloo", rank=rank, world_size=world)
    lo, hi = S.page_range(n_pages, rank, world)
    local = [f"page-{i}".encode() * (i + 1) for i in range(lo, hi)]          # variable-length byte strings, like PNGs
    out = S.gather_in_page_order(local, lo, n_pages)
    pairs = S.gather_in_page_order([(b, None if i % 3 == 0 else b[::-1]) for i, b in zip(range(lo, hi), local)], lo, n_pages)
    if rank == 0:
        assert pairs == [(b, None if i % 3 == 0 else b[::-1]) for i, b in enumerate(out)]      # (png, b64-or-None) tuples survive too
        q.put(out)
    dist.barrier()
    dist.destroy_process_group()


def test_gather_in_page_order_world2_gloo():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); port = s.getsockname()[1]; s.close()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    n_pages = 11
    procs = [ctx.Process(target=_worker, args=(r, 2, port, n_pages, q)) for r in range(2)]
    [p.start() for p in procs]
    out = q.get(timeout=120)
    [p.join(60) for p in procs]
    assert all(p.exitcode == 0 for p in procs)
    assert out == [f"page-{i}".encode() * (i + 1) for i in range(n_pages)]


def test_gather_in_page_order_same_host_shared_memory():
    """All ranks on one host (LOCAL_WORLD_SIZE == WORLD_SIZE, what torchrun sets on one box): the payload bytes go through POSIX
    shared memory, only the layout through the process group; the result is the same list."""
    s = socket.socket(); s.bind(("127.0.0.1", 0)); port = s.getsockname()[1]; s.close()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    n_pages = 13
    procs = [ctx.Process(target=_worker, args=(r, 2, port, n_pages, q, True)) for r in range(2)]
    [p.start() for p in procs]
    out = q.get(timeout=120)
    [p.join(60) for p in procs]
    assert all(p.exitcode == 0 for p in procs)
    assert out == [f"page-{i}".encode() * (i + 1) for i in range(n_pages)]


def test_keep_result_memory_is_harmless():
    """mallopt settings for result bytes: a bool comes back, and large allocations still come and go."""
    from vision_compression_project_b200 import sharding
    assert sharding.keep_result_memory() in (True, False)
    blocks = [bytes(3 << 20) for _ in range(8)]
    assert sum(len(b) for b in blocks) == 8 * (3 << 20)
    del blocks
    assert len(bytearray(5 << 20)) == 5 << 20
