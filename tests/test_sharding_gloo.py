"""N > 1 plumbing on CPU (gloo, world_size 2): the batch shards bench.py hands to each rank are disjoint, contiguous
and cover the batch; the max-over-ranks timing reduction and the gather-by-clip-index behave as the GPU arm assumes."""
import os
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def test_shard_range_partitions_any_batch():
    import bench
    for batch in (1, 2, 7, 64, 512):
        for world in (1, 2, 4, 8):
            got = []
            for r in range(world):
                lo, hi = bench.shard_range(r, world, batch)
                assert 0 <= lo <= hi <= batch
                got += list(range(lo, hi))
            assert got == list(range(batch))
    assert [bench.shard_range(r, 8, 512) for r in (0, 7)] == [(0, 64), (448, 512)]       # BASELINE configs[3]


def _worker(rank, world, port, batch, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import bench
    from neuralcodecs_b200 import synthetic
    lo, hi = bench.shard_range(rank, world, batch)
    # a stand-in for the per-clip work: a checksum that depends only on the clip index (clips are independent)
    clips = synthetic.synth_audio(hi - lo, 1000, 16000, first_clip=lo)
    local = torch.from_numpy(clips.astype(np.float64).sum(axis=1))
    sizes = [torch.zeros(1, dtype=torch.int64) for _ in range(world)]
    dist.all_gather(sizes, torch.tensor([hi - lo]))
    parts = [torch.zeros(int(s), dtype=torch.float64) for s in sizes]
    dist.all_gather(parts, local) if len(set(int(s) for s in sizes)) == 1 else None
    t = torch.tensor([10.0 + rank], dtype=torch.float64)              # per-rank step time
    dist.barrier()
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    if rank == 0:
        q.put((float(t), [int(s) for s in sizes], torch.cat(parts).numpy() if parts[0].numel() else None))
    dist.destroy_process_group()


def test_two_rank_gloo_sharding_and_max_time():
    from neuralcodecs_b200 import synthetic
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    batch, world, port = 6, 2, 29731
    procs = [ctx.Process(target=_worker, args=(r, world, port, batch, q)) for r in range(world)]
    for p in procs:
        p.start()
    t_max, sizes, gathered = q.get(timeout=120)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert t_max == 11.0 and sizes == [3, 3]
    ref = synthetic.synth_audio(batch, 1000, 16000).astype(np.float64).sum(axis=1)
    np.testing.assert_array_equal(gathered, ref)                       # same per-clip result for any shard count
