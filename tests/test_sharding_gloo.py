"""CPU, world_size 2 over gloo: the N>1 plumbing of the path -- clip sharding with no data-path collective, one
weight broadcast at init, one stats gather at exit (rmem_b200/sharding.py; reference: tools/eval.py:137-145,
managers/evaluator.py:276-295, 589-613)."""
import os
import sys

import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, q):
    sys.path.insert(0, ROOT)
    from rmem_b200 import sharding
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    sd = {"a.weight": torch.arange(12.0).view(3, 4), "b": torch.tensor([1.5, -2.0])} if rank == 0 else None
    got = sharding.broadcast_weights(sd, torch.device("cpu"), world)
    ok = torch.equal(got["a.weight"], torch.arange(12.0).view(3, 4)) and torch.equal(got["b"], torch.tensor([1.5, -2.0]))
    clips = sharding.shard_clips(7, rank, world)
    stats = sharding.gather_stats(len(clips) * 10, 0.5 * (rank + 1), torch.device("cpu"), world)
    q.put((rank, ok, clips, stats))
    dist.destroy_process_group()


def test_two_rank_sharding_over_gloo():
    world = 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + os.getpid() % 2000
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=120) for _ in range(world))
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert all(r[1] for r in res), "broadcast weights differ"
    assert res[0][2] == [0, 2, 4, 6] and res[1][2] == [1, 3, 5]
    assert sorted(res[0][2] + res[1][2]) == list(range(7))             # every clip exactly once, no overlap
    assert res[0][3] == res[1][3] == [(40, 0.5), (30, 1.0)]


def test_shard_clips_partition():
    from rmem_b200 import sharding
    for world in (1, 2, 4, 8):
        allc = sorted(c for r in range(world) for c in sharding.shard_clips(37, r, world))
        assert allc == list(range(37))
