"""CPU (gloo, world_size 2): the host-side multi-GPU plumbing — shard partitioning, the NCCL
unique-id bootstrap broadcast and max-over-ranks timing.  The collectives themselves (NCCL) are
covered by tests/test_collective_gpu.py on real GPUs; the reference likewise has no fake NCCL
(crates/burn-backend-tests/tests/tensor/distributed.rs needs >= 2 devices)."""
import os
import socket

import pytest
import torch.multiprocessing as mp

from burn_b200.distributed import shard_range


def test_shard_range_partitions_exactly():
    for n in (0, 1, 7, 64, 8192, 8193):
        for world in (1, 2, 3, 8):
            seen = []
            for r in range(world):
                lo, hi = shard_range(n, r, world)
                assert 0 <= lo <= hi <= n
                seen.extend(range(lo, hi))
            assert seen == list(range(n))
            sizes = [shard_range(n, r, world)[1] - shard_range(n, r, world)[0] for r in range(world)]
            assert max(sizes) - min(sizes) <= 1
    with pytest.raises(ValueError):
        shard_range(10, 2, 2)


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, q):
    import torch.distributed as dist
    from burn_b200.distributed import broadcast_unique_id, max_over_ranks
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    uid = broadcast_unique_id(lambda: bytes(range(128)), rank)
    slowest = max_over_ranks(10.0 + rank)
    dist.barrier()
    dist.destroy_process_group()
    q.put((rank, uid, slowest))


def test_unique_id_broadcast_and_max_timing_over_gloo():
    world = 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    results = sorted(q.get(timeout=120) for _ in range(world))
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    for rank, uid, slowest in results:
        assert uid == bytes(range(128))     # every rank holds rank 0's id
        assert slowest == 11.0              # max over ranks
