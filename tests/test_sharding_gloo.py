"""N>1 host logic on CPU: world_size-2 gloo run of the batch sharding + max-over-ranks timing that
bench.py uses (the data path itself has no collective: images are independent)."""
import os

import torch.distributed as dist
import torch.multiprocessing as mp

from rududu_image_codec_b200.sharding import job_throughput, max_over_ranks, shard


def test_shard_covers_batch_exactly_once():
    for n in (0, 1, 7, 8, 4096, 4097):
        for world in (1, 2, 3, 4, 8):
            seen = []
            for r in range(world):
                b, e = shard(n, r, world)
                assert 0 <= b <= e <= n
                seen += list(range(b, e))
            assert seen == list(range(n))
            assert max(shard(n, r, world)[1] - shard(n, r, world)[0] for r in range(world)) == (n + world - 1) // world


def _worker(rank, world, port, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    b, e = shard(13, rank, world)
    local_ms = 10.0 * (rank + 1)            # rank 1 is the slow one
    mx = max_over_ranks(local_ms)
    thr = job_throughput(e - b, local_ms)   # 13 images / 20 ms
    out.put((rank, b, e, mx, thr))
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_gloo_timing_reduction():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29700 + os.getpid() % 200
    ps = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in ps:
        p.start()
    res = sorted(q.get(timeout=120) for _ in ps)
    for p in ps:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert [(r[1], r[2]) for r in res] == [(0, 7), (7, 13)]
    for r in res:
        assert r[3] == 20.0
        assert abs(r[4] - 13 / 0.020) < 1e-6
