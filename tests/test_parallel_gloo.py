"""CPU, world_size 2, gloo: the N>1 path of the hot path (batch sharding by rank, barrier, max-over-ranks
timing, count gather).  The per-rank compute is replaced by the oracle decoder so that the shard -> decode
-> gather flow is checked end to end without a GPU."""
import os

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from openpifpaf_b200 import parallel, synth


def _worker(rank, world, port, ret):
    os.environ['MASTER_ADDR'] = '127.0.0.1'
    os.environ['MASTER_PORT'] = str(port)
    dist.init_process_group('gloo', rank=rank, world_size=world)
    try:
        from oracle import cifcaf as oc
        batch = synth.make_batch('cocokp', 5, 11, 11, 1, seed=9)       # 5 images over 2 ranks: 3 + 2
        start, stop = parallel.shard_range(5, rank, world)
        counts = []
        for b in range(start, stop):
            ann, _ = oc.decode(batch['cif'][b], 16, batch['caf'][b], 16, batch['skeleton'], 17)
            counts.append(len(ann))
        parallel.barrier()
        t = parallel.max_over_ranks(1.0 + rank)
        all_counts = parallel.gather_counts(counts)
        ret[rank] = (start, stop, t, all_counts)
    finally:
        dist.destroy_process_group()


def test_shard_decode_gather_world2():
    world = 2
    port = 29500 + (os.getpid() % 2000)
    mgr = mp.Manager()
    ret = mgr.dict()
    mp.spawn(_worker, args=(world, port, ret), nprocs=world, join=True)
    assert ret[0][:2] == (0, 3) and ret[1][:2] == (3, 5)
    assert ret[0][2] == ret[1][2] == 2.0                  # max over ranks
    assert ret[0][3] == ret[1][3] and len(ret[0][3]) == 5 and all(c == 1 for c in ret[0][3])


def test_shard_range_covers_everything():
    for n in (1, 7, 64):
        for world in (1, 2, 3, 8):
            spans = [parallel.shard_range(n, r, world) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
            assert max(s[1] - s[0] for s in spans) - min(s[1] - s[0] for s in spans) <= 1
    x = torch.arange(10).view(10, 1)
    assert torch.equal(torch.cat([parallel.shard_batch(x, r, 3) for r in range(3)]), x)
