"""CPU, world_size 2, gloo: the multi-rank host logic (shard -> search -> gather) returns the
same per-string results, in input order, as the unsharded batch.  The search itself is the
oracle here (no GPU in this container); on the GPU box the same plumbing runs the CUDA engine
(tests/test_gpu_multi.py)."""
import os
import random
import sys

import numpy as np
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, ret):
    sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
    os.environ["MASTER_ADDR"] = "127.0.0.1"; os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import oracle as O
    from libfst_b200 import shard
    f = O.Frozen.generate(O.KIND_AMBIGUOUS, 64, 12)
    rng = random.Random(5)
    strings = [bytes(rng.randint(0, 20)) for _ in range(37)]
    data = np.frombuffer(b"".join(strings), np.uint8)
    offsets = np.concatenate([[0], np.cumsum([len(s) for s in strings])]).astype(np.uint64)

    def search(d, o):
        out = []
        for i in range(len(o) - 1):
            p = O.csp_bytes(f, bytes(d[int(o[i]):int(o[i + 1])]))
            out.append((p.status, p.olabels.tolist(), p.total))
        return out

    for by_cost in (False, True):
        got, (lo, hi) = shard.search_sharded(search, data, offsets, rank, world, gather=dist.all_gather_object, by_cost=by_cost)
        want = search(data, offsets)
        assert got == want, (rank, by_cost)
        assert 0 <= lo <= hi <= len(strings)
    dist.barrier()
    dist.destroy_process_group()
    ret[rank] = True


def test_sharded_equals_unsharded_world2():
    world = 2
    port = 29500 + (os.getpid() % 2000)
    mgr = mp.Manager()
    ret = mgr.dict()
    mp.spawn(_worker, args=(world, port, ret), nprocs=world, join=True)
    assert all(ret.get(r) for r in range(world))


def test_shard_bounds_cover_everything():
    from libfst_b200 import shard
    for n in (0, 1, 7, 8, 1000):
        for w in (1, 2, 4, 8):
            b = shard.shard_bounds(n, w)
            assert b[0] == 0 and b[-1] == n and all(b[i] <= b[i + 1] for i in range(w))
            assert max(b[i + 1] - b[i] for i in range(w)) - min(b[i + 1] - b[i] for i in range(w)) <= 1
    lens = np.array([11, 251, 19, 224, 33, 192, 64, 160, 96, 128] * 10)
    b = shard.shard_by_cost(lens, 4)
    assert b[0] == 0 and b[-1] == len(lens) and all(b[i] <= b[i + 1] for i in range(4))
