"""N > 1 host logic on CPU with a world-size-2 gloo group: the batch shards with no data-path collective;
only byte strings and timings are exchanged (compressai/utils/sharding.py, used by bench.py)."""
import os
import socket

import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from compressai.utils.sharding import gather_strings, max_over_ranks, scatter_strings, shard_range


def test_shard_range_partitions_exactly():
    for n in (0, 1, 7, 8, 63, 64, 65):
        for world in (1, 2, 3, 4, 8):
            spans = [shard_range(n, r, world) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
            sizes = [hi - lo for lo, hi in spans]
            assert max(sizes) - min(sizes) <= 1
    assert [shard_range(64, r, 8) for r in (0, 7)] == [(0, 8), (56, 64)]  # BASELINE configs[2] at 8 GPUs


def _worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        n = 5
        lo, hi = shard_range(n, rank, world)
        mine = [[b"y%d" % i * (i + 1) for i in range(lo, hi)], [b"z%d" % i for i in range(lo, hi)]]
        allstr = gather_strings(mine, dst=0)
        slow = max_over_ranks(10.0 + rank)
        back = scatter_strings(allstr, n, src=0)
        dist.barrier()
        q.put((rank, allstr, slow, back == mine))
    finally:
        dist.destroy_process_group()


def test_gather_scatter_and_timing_with_gloo_world2():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=120) for _ in procs)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    (r0, all0, slow0, ok0), (r1, all1, slow1, ok1) = res
    assert all1 is None and ok0 and ok1
    assert all0 == [[b"y%d" % i * (i + 1) for i in range(5)], [b"z%d" % i for i in range(5)]]
    assert slow0 == slow1 == 11.0
