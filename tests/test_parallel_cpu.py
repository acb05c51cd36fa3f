"""Host-side multi-GPU logic on CPU with the gloo backend, world_size 2."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, ret):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from mebt_b200 import parallel as P
    # data-parallel gradient buckets: 6 blocks of 10 params, head 4, embeddings 7
    block_slices = [(i * 10, (i + 1) * 10) for i in range(6)]
    chunks = [(0, 2), (2, 4), (4, 6)]
    head, emb = (60, 64), (64, 71)
    buckets = P.bucket_slices(block_slices, chunks, head, emb)
    assert buckets == [(40, 60), (60, 64), (20, 40), (0, 20), (64, 71)]
    g = torch.Generator().manual_seed(100 + rank)
    flat = torch.randn(71, generator=g)
    mine = flat.clone()
    for lo, hi in buckets:
        P.allreduce_mean_(flat[lo:hi])
    other = torch.randn(71, generator=torch.Generator().manual_seed(100 + (1 - rank)))
    assert torch.allclose(flat, (mine + other) / 2)
    # after the all-reduce every rank holds bit-identical gradients
    ref = flat.clone()
    dist.broadcast(ref, 0)
    assert torch.equal(ref, flat)
    # sharded exchange (fused flat AdamW): reduce-scatter -> update of the own shard -> all-gather of the result
    g2 = torch.Generator().manual_seed(200 + rank)
    grads = torch.randn(64, generator=g2)
    other_g = torch.randn(64, generator=torch.Generator().manual_seed(200 + (1 - rank)))
    params = torch.arange(64, dtype=torch.float32)
    for lo, hi in [(40, 60), (60, 64), (20, 40), (0, 20)]:
        a, b = P.shard_bounds(lo, hi, rank, world)
        shard = P.reduce_scatter_mean_(grads[lo:hi])
        assert shard.data_ptr() == grads[a:b].data_ptr()
        assert torch.allclose(grads[a:b], (grads.new_tensor(0) + (other_g[a:b] + torch.randn(64, generator=torch.Generator().manual_seed(200 + rank))[a:b]) / 2))
        params[a:b] -= 0.5 * grads[a:b]                       # "optimizer" on the shard only
        P.all_gather_shards_(params[lo:hi])
    mean = (torch.randn(64, generator=torch.Generator().manual_seed(200)) + torch.randn(64, generator=torch.Generator().manual_seed(201))) / 2
    assert torch.allclose(params, torch.arange(64, dtype=torch.float32) - 0.5 * mean)
    chk = params.clone()
    dist.broadcast(chk, 0)
    assert torch.equal(chk, params)                           # every rank ends with the same parameters
    with pytest.raises(ValueError):
        P.shard_bounds(0, 7, rank, world)
    # sampling: videos sharded by batch, per-rank RNG streams, results gathered on rank 0 for output only
    shard = P.shard_range(7, rank, world)
    ids = torch.full((len(shard), 4), rank, dtype=torch.long)
    sizes = [len(P.shard_range(7, r, world)) for r in range(world)]
    assert sum(sizes) == 7 and max(sizes) - min(sizes) <= 1
    assert P.rank_seed(1000, rank) == 1000 + rank
    if sizes[0] == sizes[1]:
        out = P.gather_ids(ids)
        if rank == 0:
            assert [int(o[0, 0]) for o in out] == [0, 1]
    dist.destroy_process_group()
    ret[rank] = True


def test_bucketed_allreduce_and_video_sharding_gloo():
    world = 2
    port = _free_port()
    with mp.Manager() as m:
        ret = m.dict()
        mp.spawn(_worker, args=(world, port, ret), nprocs=world, join=True)
        assert all(ret.get(r) for r in range(world))


def test_shard_range_partitions():
    from mebt_b200.parallel import shard_range
    for n in (0, 1, 7, 8, 33):
        for world in (1, 2, 4, 8):
            covered = [i for r in range(world) for i in shard_range(n, r, world)]
            assert covered == list(range(n))
